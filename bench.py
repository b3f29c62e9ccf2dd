#!/usr/bin/env python
"""Benchmark of the RBM CD-k hot path on B200 (BASELINE.json metric: RBM CD-k training samples/sec).

    python bench.py --gpus N --steps K --warmup W            # the engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Bernoulli RBM
4096 -> 4096, CD-10, minibatch 4096 rows per GPU (weak scaling), bf16 contractions, synthetic binarised
data.  A "step" is one minibatch: 21 projections with fused sigmoid/Philox/threshold epilogues, the dW
contraction, the all-reduce (N > 1) and the fused update.  `value` is measured with every minibatch
already resident in HBM (K distinct minibatches, far larger than L2, replayed as one CUDA graph per
step); `e2e` goes through the reference-facing call with float32 HOST (pinned) minibatches, the
host->device copy and the device->host read of the step's statistic inside the timed region.

One JSON line on stdout (rank 0).  Under torchrun each rank drives one GPU; time is the max over ranks
of CUDA-event time on the engine's stream.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (V, H, batch per GPU, k, dtype, bernoulli prob of a 1)
    "c3": dict(V=4096, H=4096, B=4096, k=10, dtype="bf16", q=0.5,
               desc="Bernoulli RBM 4096->4096, CD-10, batch 4096 per GPU, bf16 (BASELINE.json configs[2])"),
    "c3f32": dict(V=4096, H=4096, B=4096, k=10, dtype="f32", q=0.5,
                  desc="Bernoulli RBM 4096->4096, CD-10, batch 4096 per GPU, float32-grade contractions (three bf16 terms)"),
    "c1": dict(V=784, H=500, B=128, k=1, dtype="bf16", q=0.1307,
               desc="Bernoulli RBM 784->500, CD-1, batch 128 (BASELINE.json configs[0])"),
    "c1f32": dict(V=784, H=500, B=128, k=1, dtype="f32", q=0.1307,
                  desc="Bernoulli RBM 784->500, CD-1, batch 128, float32-grade contractions (BASELINE.json configs[0])"),
    "c4": dict(V=16384, H=8192, B=1024, k=1, dtype="bf16", q=0.5, persistent=True,
               desc="PCD RBM 16384->8192, 1024 rows + 1024 persistent chains per GPU (8192 over 8 GPUs), bf16 "
                    "(BASELINE.json configs[3])"),
    # handled by run_dbn / run_infer
    "c2": dict(V=784, H=500, B=256, k=1, dtype="bf16", q=0.1307, layers=[784, 500, 500, 2000],
               desc="DBN 784-500-500-2000 greedy layer-wise CD-1 pretraining, batch 256 (BASELINE.json configs[1])"),
    "c5": dict(V=4096, H=4096, B=0, k=0, dtype="bf16", q=0.5,
               desc="RBM transform / free-energy inference sweep, 1K..1M rows, 4096->4096 (BASELINE.json configs[4])"),
}
METRIC = "RBM CD-k training samples/sec"
UNIT = "samples/s"


def flops_per_sample(V, H, k):
    return (2 * k + 3) * 2 * V * H  # (2k+1) projections + 2 outer products


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the benchmark runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = []
        for t, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                rows.append((t, float(f[0]), float(f[1]), float(f[2]), f[3:7]))
            except ValueError:
                continue
        inside = [r for r in rows if any(a - 0.02 <= r[0] <= b + 0.02 for a, b in windows)]
        used = inside if len(inside) >= 2 else [r for r in rows if r[3] > 300.0] or rows
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(r[1] for r in used)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in used for i in range(4) if r[4][i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": used[0][2], "reasons": reasons,
                "power_w_max": max(r[3] for r in used), "samples": len(used),
                "window": "timed regions" if used is inside else "under load (timed region shorter than the sampling period)"}


def cpu_threads_setup():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference's path "with all the host
    threads it can use".  Called before numpy (OpenBLAS) is first imported."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    return n


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([int(i.get("num_threads", 1)) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
    except Exception:  # noqa: BLE001
        return None


def cpu_step(orc, x, k, rng, schedule):
    """One CD-k minibatch of the oracle port in float32, drawing its uniforms as TF would.
    schedule 'fused': one chain, W / b / c from the same statistics (the engine's schedule; 2k+3 contractions).
    schedule 'reference': what ku/ebm/rbm.py:214-233 executes per minibatch - three sequential single-parameter
    runs with fresh draws, two free energies and a fourth chain for the printed score (6k+8 contractions; 14 at
    k = 1)."""
    import numpy as np

    rows = x.shape[0]

    def draws():
        return ([rng.random((rows, orc.H), dtype=np.float32) for _ in range(k)],
                [None] + [rng.random((rows, orc.V), dtype=np.float32) for _ in range(k)])

    if schedule == "fused":
        u_h, u_v = draws()
        orc.fused_step(x, u_h, u_v, lr=1e-3, k=k, scale=1.0 / rows)
    else:
        d = [draws() for _ in range(3)]
        d.append((rng.random((rows, orc.H), dtype=np.float32), rng.random((rows, orc.V), dtype=np.float32)))
        orc.reference_step(x, d, lr=1e-3, k=k, scale=1.0 / rows)


def cpu_arm(cfg, steps, warmup, rows, schedule="fused"):
    """The reference's CPU path: the oracle port of ku/ebm/rbm.py (TensorFlow is not installable),
    float32 numpy/BLAS on all host cores, `rows` rows per step (a bounded sample of the minibatch)."""
    import numpy as np

    from oracle import cd_oracle as O

    rng = np.random.default_rng(7)
    W, b, c = O.OracleRBM.init_params(cfg["V"], cfg["H"], seed=0)
    orc = O.OracleRBM(W, b, c, compute="f32")
    x = (np.random.default_rng(1234).random((rows, cfg["V"])) < cfg["q"]).astype(np.float32)
    for _ in range(warmup):
        cpu_step(orc, x, cfg["k"], rng, schedule)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(orc, x, cfg["k"], rng, schedule)
    dt = time.perf_counter() - t0
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count()
    return dict(value=rows * steps / dt, seconds=dt, cores=cores, rows=rows, steps=steps, schedule=schedule,
                blas_threads=blas_threads())


def cpu_baseline_obj(cfg, steps, rows, B):
    """Both CPU figures BASELINE.md section 4 promises: `value` is the reference's own schedule (rbm.py:214-233), the
    fused one rides along so that the ratio is not inflated by the reference's redundant passes."""
    k = cfg["k"]
    ref = cpu_arm(cfg, steps, 1, rows, "reference")
    fused = cpu_arm(cfg, steps, 1, rows, "fused")
    return {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "port",
            "blas_threads": ref["blas_threads"],
            "sample": "%d steps of %d rows (of the %d-row minibatch), float32 numpy/BLAS oracle port of ku/ebm/rbm.py, "
                      "the reference's schedule (rbm.py:214-233: runs A, B, C with fresh draws, two free energies, score "
                      "chain = %d contractions per minibatch at CD-%d), %.1f s" % (ref["steps"], rows, B, 6 * k + 8, k,
                                                                                   ref["seconds"]),
            "fused_schedule": {"value": fused["value"], "unit": UNIT,
                               "sample": "same rows, one chain with W / b / c from the same statistics (%d contractions, "
                                         "the schedule the GPU arm runs), %.1f s" % (2 * k + 3, fused["seconds"])}}


def parse_cpulist(txt):
    """'0-3,8,10-11' (sysfs cpulist format) -> {0, 1, 2, 3, 8, 10, 11}"""
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_near_gpu(local_rank):
    """One process per GPU on a multi-socket host: keep this rank's host threads - and with them (first touch) its
    page-locked staging memory - on the CPUs local to its GPU, so that eight ranks streaming minibatches do not all pull
    through one socket.  Returns the CPU list it bound to, or None (single NUMA node, restricted cpuset, no sysfs ...).
    KUCD_BENCH_NUMA=0 turns it off."""
    if os.environ.get("KUCD_BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(path) as f:
            txt = f.read().strip()
        cpus = parse_cpulist(txt)
        cur = os.sched_getaffinity(0)
        keep = cpus & cur
        if len(keep) < 4 or keep == cur:
            return None
        os.sched_setaffinity(0, keep)
        return txt
    except Exception:  # noqa: BLE001 - a placement hint, never an error
        return None


def _events(ctx, local_rank):
    import torch

    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local_rank))
    return ext, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run_dbn(cfg, ctx, compute, steps, warmup, config):
    """C2: greedy layer-wise pretraining of a 784-500-500-2000 DBN (dbn.py:51-55): per layer one epoch of CD-1
    over N = steps * 256 rows, then the full-data transform that feeds the next layer - all on the GPU.  One
    whole pass is the warm-up, a second whole pass (fresh models; includes each layer's graph capture and the
    allocation of the inter-layer data sets, as a user's DBN.fit does) is timed."""
    import numpy as np
    import torch

    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    dims, B = cfg["layers"], cfg["B"]
    N = steps * B
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    X = (torch.rand((N, dims[0]), device="cuda", generator=gen) < cfg["q"]).to(torch.uint8)
    ds0 = Dataset.from_array(ctx, X, compute)
    hp = Machine.hparams(lr=1e-3, k=1)
    prng = np.random.default_rng(0)

    def one_pass():
        ms, cur, made = [], ds0, []
        for i in range(len(dims) - 1):
            m = Machine(ctx, dims[i], dims[i + 1], L.MODE_VISIBLE_BERNOULLI, compute, seed=42 + i)
            m.set_params(prng.uniform(-0.05, 0.05, (dims[i], dims[i + 1])).astype(np.float32),
                         prng.uniform(-0.05, 0.05, dims[i]).astype(np.float32),
                         prng.uniform(-0.05, 0.05, dims[i + 1]).astype(np.float32))
            ms.append(m)
        ctx.sync()
        ext, e0, e1 = _events(ctx, ctx.device)
        ctx.timings(reset=True)
        tw0 = time.time()
        e0.record(ext)
        for i, m in enumerate(ms):
            m.fit_epoch(cur, B, hp, want_stats=False)
            if i + 1 < len(ms):
                cur = m.transform_dataset(cur)
                made.append(cur)
        e1.record(ext)
        ctx.sync()
        torch.cuda.synchronize()
        tw1 = time.time()
        t = ctx.timings()
        for d in made:
            d.close()
        for m in ms:
            m.close()
        return e0.elapsed_time(e1), t, (tw0, tw1)

    one_pass()
    ms_total, t, win = one_pass()
    flop = sum((2 * 1 + 3) * 2 * dims[i] * dims[i + 1] for i in range(len(dims) - 1)) + \
        sum(2 * dims[i] * dims[i + 1] for i in range(len(dims) - 2))
    config = dict(config, layers=dims, rows=N, batch_per_gpu=B, k=1)
    return {"metric": METRIC, "value": N / (ms_total * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": cfg["dtype"], "data": "synthetic binarised (Bernoulli %.4g)" % cfg["q"], "config": config,
            "gpu_launches": int(t["graph_kernel_launches"] + t["gemm_launches"] + t["aux_launches"]),
            "note": "a sample passes through all three layers; latency-bound (0.5-2.6 GFLOP per step)",
            "tflops": flop * N / (ms_total * 1e-3) / 1e12, "_windows": [win]}


def run_infer(cfg, ctx, compute, steps, warmup, config):
    """C5: transform (sampled hidden states, rbm.py:88) and free energy (rbm.py:97) over n rows resident in HBM."""
    import numpy as np
    import torch

    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    V, H = cfg["V"], cfg["H"]
    m = Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, compute, seed=42)
    prng = np.random.default_rng(0)
    m.set_params(prng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), prng.uniform(-0.05, 0.05, V).astype(np.float32),
                 prng.uniform(-0.05, 0.05, H).astype(np.float32))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    pk = peaks()
    sweep, windows = [], []
    for logn in (10, 12, 14, 16, 18, 20):
        n = 1 << logn
        X = torch.empty((n, V), dtype=torch.uint8, device="cuda")
        for r0 in range(0, n, 65536):
            r1 = min(n, r0 + 65536)
            X[r0:r1] = (torch.rand((r1 - r0, V), device="cuda", generator=gen) < cfg["q"]).to(torch.uint8)
        ds = Dataset.from_array(ctx, X, compute)
        reps = max(3, min(50, (1 << 22) // n))
        for _ in range(3):
            m.transform_dataset(ds).close()
        ctx.sync()
        ext, e0, e1 = _events(ctx, ctx.device)
        outs = []
        tw0 = time.time()
        e0.record(ext)
        for _ in range(reps):
            outs.append(m.transform_dataset(ds))
        e1.record(ext)
        ctx.sync()
        torch.cuda.synchronize()
        t_tr = e0.elapsed_time(e1) / reps
        for o in outs:
            o.close()
        nfe = min(n, 1 << 18)
        Xfe = X[:nfe]
        m.free_energy(Xfe)
        e0.record(ext)
        for _ in range(3):
            m.free_energy(Xfe)
        e1.record(ext)
        ctx.sync()
        torch.cuda.synchronize()
        t_fe = e0.elapsed_time(e1) / 3
        windows.append((tw0, time.time()))
        sweep.append({"rows": n, "transform_ms": t_tr, "transform_rows_per_s": n / (t_tr * 1e-3),
                      "transform_tflops": 2.0 * n * V * H / (t_tr * 1e-3) / 1e12,
                      "transform_frac_of_sustained": 2.0 * n * V * H / (t_tr * 1e-3) / 1e12 / pk["sustained"],
                      "free_energy_rows": nfe, "free_energy_rows_per_s": nfe / (t_fe * 1e-3)})
        ds.close()
        del X
    best = sweep[-1]
    return {"metric": "RBM transform rows/sec", "value": best["transform_rows_per_s"], "unit": "rows/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": best["transform_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"],
            "data": "synthetic binarised (Bernoulli %.4g)" % cfg["q"], "config": config, "sweep": sweep,
            "note": "transform: resident data set in, resident data set out, one launch (allocation of the output "
                    "included); free energy: device tensor in, device vector out, 32768-row chunks", "_windows": windows}


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line there.
    Everything else goes to stderr: keep the real stdout aside and point fd 1 at fd 2."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out = _claim_stdout()
    try:
        return _main(out)
    finally:
        out.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-f32-grade", action="store_true", help="skip the float32-grade figure of the c3 line (profiling runs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's minibatch PER GPU (default, what the metric is quoted on); strong: that "
                         "minibatch is the GLOBAL one, split by rows over the GPUs (SURVEY.md 8d: report both - the weak "
                         "line carries the strong figure as `strong_scaling` when N > 1)")
    ap.add_argument("--batch", type=int, default=0,
                    help="rows per GPU instead of the workload's (e.g. 512: the per-GPU share of C3 under strong scaling over "
                         "8 GPUs, measured on fewer); the line says so in config.workload")
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        cfg["B"] = args.batch
        cfg["desc"] += " [--batch %d rows per GPU instead of the workload's]" % args.batch
    V, H, B, k = cfg["V"], cfg["H"], cfg["B"], cfg["k"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    B_global = B * world   # weak scaling: the workload's minibatch on every GPU
    if args.scaling == "strong" and world > 1 and B:
        if B % (world * 128):
            raise SystemExit("--scaling strong: the %d-row minibatch does not split into %d shards of whole 128-row blocks"
                             % (B, world))
        B_global, B = B, B // world
    config = {"workload": cfg["desc"], "n_visible": V, "n_hidden": H, "batch_per_gpu": B, "global_batch": B_global,
              "k": k, "schedule": "fused single chain (W, b, c from the same statistics)",
              "parallelism": "dp%d" % world,
              "l2": "every step reads a different minibatch; the resident data set is far larger than L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu_threads_setup()
        rows = 256 if args.workload in ("c3", "c3f32") else min(B, 1024)
        n = max(1, min(steps, 20))
        ref = cpu_arm(cfg, n, min(warmup, 1), rows, "reference")
        fused = cpu_arm(cfg, n, min(warmup, 1), rows, "fused")
        sample = ("%d steps of %d rows (a bounded sample of the %d-row global minibatch: rows are independent, so CPU "
                  "samples/s does not depend on it), float32 numpy/BLAS oracle port of ku/ebm/rbm.py on %d host threads, "
                  "the reference's own schedule rbm.py:214-233 (three sequential single-parameter runs, two free energies, "
                  "score chain: %d contractions per minibatch at CD-%d)" % (n, rows, B * world, ref["blas_threads"] or
                                                                            ref["cores"], 6 * k + 8, k))
        line = {"metric": METRIC, "value": ref["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": n, "warmup": min(warmup, 1), "ms_per_step": 1e3 * ref["seconds"] / n,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "port",
                                 "blas_threads": ref["blas_threads"], "sample": sample,
                                 "fused_schedule": {"value": fused["value"], "unit": UNIT,
                                                    "sample": "same rows, the GPU arm's schedule (one chain, %d "
                                                              "contractions)" % (2 * k + 3)}},
                "e2e": {"value": ref["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "one host, all of its threads, whatever --gpus says: the CPU path does not scale with GPUs; the "
                        "reference computes in float32, the GPU arm's default workload in bf16 (its float32-grade figure "
                        "is the `f32_grade` key of the GPU line)",
                "gpu_launches": 0}
        print(json.dumps(line), file=out)
        return 0

    switches = {k_: v_ for k_, v_ in sorted(os.environ.items()) if k_.startswith("KUCD_")}
    if switches:  # a line measured off the defaults says so ("Switches")
        config["switches"] = switches

    import numpy as np
    import torch

    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    torch.cuda.set_device(local_rank)
    host_cpus = bind_near_gpu(local_rank) if world > 1 else None  # before any host buffer of this rank exists
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Context(device=local_rank, seed=42)
    if world > 1:
        ctx.join_group(rank, world)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    compute = L.COMPUTE_BF16 if cfg["dtype"] == "bf16" else L.COMPUTE_F32X3
    if args.workload in ("c2", "c5"):
        fn = run_dbn if args.workload == "c2" else run_infer
        line = fn(cfg, ctx, compute, steps, warmup, config)
        line["clocks"] = sampler.stop(line.pop("_windows")) if sampler is not None else None
        print(json.dumps(line), file=out)
        return 0
    m = Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, compute, seed=42)
    prng = np.random.default_rng(0)
    m.set_params(prng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), prng.uniform(-0.05, 0.05, V).astype(np.float32),
                 prng.uniform(-0.05, 0.05, H).astype(np.float32))

    # K distinct synthetic binarised minibatches per rank, resident in HBM as the engine's operand planes
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    n_batches = max(steps, 1)
    X = torch.empty((n_batches * B, V), dtype=torch.uint8, device="cuda")
    for i in range(n_batches):
        X[i * B:(i + 1) * B] = (torch.rand((B, V), device="cuda", generator=gen) < cfg["q"]).to(torch.uint8)
    torch.cuda.synchronize()
    ds = Dataset.from_array(ctx, X, compute)
    persistent = bool(cfg.get("persistent"))
    if persistent:
        m.set_chains((torch.rand((B, V), device="cuda", generator=gen) < 0.5).to(torch.uint8))
    hp = Machine.hparams(lr=1e-3, k=k, normalize=True, persistent=persistent)
    row0 = rank * B

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- warm-up (captures the step graph), then exactly K timed steps -----------------------------
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local_rank))

    def timed_region(b_loc, g_row0):
        """W warm-up steps, then K timed steps of b_loc rows per rank; CUDA events on the engine stream, max over ranks."""
        done = 0
        while done < warmup:
            n = min(warmup - done, n_batches)
            m.fit_range(ds, b_loc, hp, 0, n, global_row0=g_row0)
            done += n
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.timings(reset=True)
        tw0 = time.time()
        e0.record(ext)
        m.fit_range(ds, b_loc, hp, 0, steps, global_row0=g_row0)
        e1.record(ext)
        ctx.sync()
        torch.cuda.synchronize()
        tw1 = time.time()
        ms_ = e0.elapsed_time(e1)
        tm_ = ctx.timings()
        if dist is not None:
            t = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, tm_, (tw0, tw1)

    def exchange_of(tm_):
        if world == 1:
            return "none (one GPU)"
        if tm_.get("unit_steps"):
            return ("unit-sharded: every rank computes its slice of the units for the whole global minibatch, 0/1 states "
                    "cross NVLink as bits after each projection, dW and the update stay local (%d bit exchanges per step)"
                    % (tm_["unit_exchanges"] // max(tm_["unit_steps"], 1)))
        if tm_["fused_reduce_steps"]:
            return "fused: dW rows stored into their owners' memory by the contraction epilogue (NVLink), flag barriers"
        return "ncclAllReduce(dW | db | dc)"

    ms, tm, win = timed_region(B, row0)
    launches = tm["graph_kernel_launches"] + 1  # + the step-state initialisation kernel
    value = world * steps * B / (ms * 1e-3)
    windows = [win]
    config["exchange"] = exchange_of(tm)

    # ---- strong scaling beside the weak line (N > 1): the workload's minibatch as the GLOBAL one, split by rows ----
    strong = None
    if args.scaling == "weak" and world > 1 and B % (world * 128) == 0 and not persistent:
        bs = B // world
        ms_s, tm_s, win_s = timed_region(bs, rank * bs)
        strong = {"value": steps * B / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / steps, "global_batch": B,
                  "batch_per_gpu": bs, "exchange": exchange_of(tm_s), "steps": steps,
                  "note": "same minibatch as one GPU's, split by rows: per-GPU compute shrinks by N while the dW exchange "
                          "stays %d MiB" % (V * H * 4 >> 20)}
        windows.append(win_s)

    # ---- end to end: float32 HOST minibatches through the reference-facing path (what RBM.fit(V) with a numpy
    # V runs for one epoch): kucd_rbm_fit_host copies minibatch i+1 from pinned host memory while minibatch i
    # trains, and reads every step's statistic back; all of that is inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        # long steps: 40 minibatches bound the pass; short steps (C1, C2): as many as the timed region, so that the
        # fixed cost of one call (staging buffers, the first copy that nothing overlaps) weighs as it does in a real fit
        n_e2e = max(3, min(steps, 40)) if B * V >= (1 << 20) else max(3, steps)
        host = torch.empty((n_e2e * B, V), dtype=torch.float32).pin_memory()
        for i in range(n_e2e):
            host[i * B:(i + 1) * B].copy_(X[(i % n_batches) * B:((i % n_batches) + 1) * B].to(torch.float32))
        hp_e = Machine.hparams(lr=1e-3, k=k, normalize=True, persistent=persistent)
        m.fit_host(host[:2 * B], B, hp_e, global_row0=row0)  # warm-up: staging buffers, pinned result buffer
        barrier()
        ctx.timings(reset=True)
        tw0 = time.time()
        t0 = time.perf_counter()
        st = m.fit_host(host, B, hp_e, global_row0=row0)   # returns after the last statistic has been read back
        dt = time.perf_counter() - t0
        tw1 = time.time()
        te = ctx.timings()
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_e2e * B / dt, "unit": UNIT, "h2d_bytes_per_step": te["h2d_bytes"] // n_e2e,
               "d2h_bytes_per_step": te["d2h_bytes"] // n_e2e, "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e,
               "input": "float32 pinned host matrix, one pass of kucd_rbm_fit_host (copy of minibatch i+1 overlapped "
                        "with minibatch i); result read per step: recon_err; wall clock around the call",
               "last_recon_err": float(st["step_recon_err"][-1])}
        if host_cpus is not None:
            e2e["host_cpus"] = "rank 0 bound to the CPUs local to its GPU (%s); every rank does the same" % host_cpus
        windows.append((tw0, tw1))
        del host
    if e2e is not None:
        # The same pass with the binarised data handed over as uint8 (a quarter of the bytes) and as packed bits (1/32 of
        # them; data.PackedBits, include/kucd.h bits = 1).  Informational: a failure here must not cost the line.
        def timed_pass(arr):
            m.fit_host(arr[:2 * B], B, hp_e, global_row0=row0)
            barrier()
            t0 = time.perf_counter()
            m.fit_host(arr, B, hp_e, global_row0=row0)
            dt_ = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt_], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt_ = float(t.item())
            return dt_

        try:
            host8 = torch.empty((n_e2e * B, V), dtype=torch.uint8).pin_memory()
            for i in range(n_e2e):
                host8[i * B:(i + 1) * B].copy_(X[(i % n_batches) * B:((i % n_batches) + 1) * B])
            dt8 = timed_pass(host8)
            e2e["uint8_input"] = {"value": world * n_e2e * B / dt8, "unit": UNIT, "h2d_bytes_per_step": B * V,
                                  "ms_per_step": 1e3 * dt8 / n_e2e}
            del host8
        except Exception as exc:  # noqa: BLE001
            e2e["uint8_input"] = {"error": repr(exc)}
        try:
            from keras_unsupervised_b200.data import PackedBits

            nb = (V + 7) // 8
            hostb = torch.empty((n_e2e * B, nb), dtype=torch.uint8).pin_memory()
            weights = (2 ** torch.arange(8, device="cuda", dtype=torch.int32)).view(1, 1, 8)
            for i in range(n_e2e):
                blk = X[(i % n_batches) * B:((i % n_batches) + 1) * B].to(torch.int32)
                if V % 8:
                    blk = torch.nn.functional.pad(blk, (0, nb * 8 - V))
                hostb[i * B:(i + 1) * B].copy_((blk.view(B, nb, 8) * weights).sum(dim=2).to(torch.uint8))
            dtb = timed_pass(PackedBits(hostb, V))
            e2e["packed_bits_input"] = {"value": world * n_e2e * B / dtb, "unit": UNIT, "h2d_bytes_per_step": B * nb,
                                        "ms_per_step": 1e3 * dtb / n_e2e}
            del hostb
        except Exception as exc:  # noqa: BLE001
            e2e["packed_bits_input"] = {"error": repr(exc)}

    # ---- roofline of the dominant kernel: per-launch CUDA events on the engine stream -------------
    roof = None
    pk = peaks()
    if True:  # every rank runs these steps (they all-reduce); rank 0 reports
        barrier()
        ctx.set_profile(True)
        ctx.timings(reset=True)
        hp_d = Machine.hparams(lr=1e-3, k=k, normalize=True, persistent=persistent)
        n_prof = 3
        tw0 = time.time()
        for i in range(n_prof):
            m.cd_step(X[(i % n_batches) * B:((i % n_batches) + 1) * B], hp_d, global_row0=row0)
        tp = ctx.timings()
        tw1 = time.time()
        ctx.set_profile(False)
        windows.append((tw0, tw1))
        if tp["proj_timed"] and rank == 0:
            n_proj = 2 * k + 1 + (1 if persistent else 0)
            chain = tp["chain_launches"] > 0
            if chain:
                # all projections of a minibatch are ONE launch of chain_kernel: that launch is the unit
                n_timed = tp["chain_launches"]
                with_dw = tp["chain_dw_launches"] > 0
                flop = (n_proj + (2 if with_dw else 0)) * 2.0 * B * V * H
                kernel = "chain_kernel (the %d projections of a CD-%d minibatch%s, fused epilogues, one persistent launch)" % (
                    n_proj, k, " + the two outer products of dW" if with_dw else "")
            else:
                n_timed = tp["proj_timed"]  # with the two-chain schedule each projection is two row-half launches
                flop = n_prof * n_proj * 2.0 * B * V * H / n_timed
                kernel = "gemm_bf16_kernel<sample epilogue> (v.W+c / h.W^T+b projection)"
            per_launch_ms = tp["proj_ms"] / n_timed
            ach = flop / (per_launch_ms * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tpath):
                with open(tpath) as f:
                    traffic = json.load(f).get(args.workload + ("_chain" if chain else ""))
            step_tf = flops_per_sample(V, H, k) * B / (ms * 1e-3 / steps) / 1e12
            # Which measured peak applies (MEASURED_PEAKS.json: `bf16_tflops` is the best of 10 single matmuls - a burst at
            # full clocks -, `bf16_tflops_sustained` 4 s of back-to-back matmuls under the power cap)?  A timed region of
            # less than a second never reaches the steady power-capped state: burst.  Both fractions are printed.
            regime = "burst" if ms * 1e-3 < 1.0 else "sustained"
            peak = pk[regime]
            roof = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "regime": regime,
                    "frac_of_burst_peak": ach / pk["burst"], "frac_of_sustained_peak": ach / pk["sustained"],
                    "peak_source": pk["source"] + ": %s figure, because the timed region lasted %.3f s (burst below 1 s, "
                                                  "sustained above)" % (regime, ms * 1e-3),
                    "measured": "CUDA events on the engine stream around each launch of %d further steps issued directly "
                                "(same process, same data, right after the timed region; a graph replay cannot carry "
                                "per-kernel events)" % n_prof,
                    "traffic": traffic, "launch_ms": per_launch_ms, "launches_timed": n_timed,
                    "projections_per_launch": n_proj if chain else None, "flop_per_launch": flop,
                    "dw_launch_ms": (tp["dw_ms"] / tp["dw_timed"]) if tp["dw_timed"] else None,
                    "step_breakdown_ms": {
                        "what": "device time per directly launched step, by kernel class (CUDA events around each class; "
                                "the replayed step of the timed region runs the same kernels)",
                        "projections": tp["proj_ms"] / n_prof, "dW": tp["dw_ms"] / n_prof,
                        "bit_exchanges": tp.get("xchg_ms", 0.0) / n_prof, "n_bit_exchanges": tp.get("xchg_timed", 0) // n_prof,
                        "update_and_closing_barrier": tp.get("upd_ms", 0.0) / n_prof},
                    "step_tflops": step_tf, "step_frac": step_tf / peak,
                    "step_frac_of_burst_peak": step_tf / pk["burst"],
                    "step_frac_of_sustained_peak": step_tf / pk["sustained"]}

    # ---- the reference's own precision beside the bf16 headline (rbm.py:39 K.floatx() = float32): the same workload
    # with float32-grade contractions (three bf16 terms per operand), a few steps -----------------------------------
    f32_grade = None
    if args.workload == "c3" and world == 1 and not args.no_f32_grade:
        try:
            barrier()
            m32 = Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_F32X3, seed=42)
            m32.set_params(*m.get_params())
            n32 = max(3, min(steps, 10))
            ds32 = Dataset.from_array(ctx, X[:n32 * B], L.COMPUTE_F32X3)
            m32.fit_range(ds32, B, hp, 0, min(3, n32), global_row0=row0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tw0 = time.time()
            e0.record(ext)
            m32.fit_range(ds32, B, hp, 0, n32, global_row0=row0)
            e1.record(ext)
            ctx.sync()
            torch.cuda.synchronize()
            ms32 = e0.elapsed_time(e1)
            windows.append((tw0, time.time()))
            f32_grade = {"value": n32 * B / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / n32, "steps": n32,
                         "dtype": "f32 (three bf16 terms per operand, fp32 accumulation: 1e-5 relative to float64)",
                         "tflops": flops_per_sample(V, H, k) * n32 * B / (ms32 * 1e-3) / 1e12}
            ds32.close()
            m32.close()
        except Exception as exc:  # noqa: BLE001 - informational: must not cost the line
            f32_grade = {"error": repr(exc)}
    clocks = sampler.stop(windows) if sampler is not None else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        big = args.workload in ("c3", "c3f32", "c4")
        cpu = cpu_baseline_obj(cfg, 3 if big else 20, 512 if big else B, B)

    if dist is not None:
        dist.barrier()
    if rank == 0:
        tf_gpu = flops_per_sample(V, H, k) * value / world / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": cfg["dtype"], "data": "synthetic binarised (Bernoulli %.4g), U(-0.05,0.05) weights" % cfg["q"],
                "config": config, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof,
                "cpu_baseline": cpu,
                "per_gpu": {"samples_per_s": value / world, "tflops": tf_gpu,
                            "frac_of_burst_bf16_peak": tf_gpu / pk["burst"],
                            "frac_of_sustained_bf16_peak": tf_gpu / pk["sustained"]}}
        if strong is not None:
            line["strong_scaling"] = strong
        if f32_grade is not None:
            line["f32_grade"] = f32_grade
        print(json.dumps(line), file=out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
