"""The engine's HOST logic on the CPU: csrc/kucd.cu linked against a fake CUDA runtime (tools/dryrun/fake_cudart.cpp) that
records kernel launches instead of running them and checks every copy, memset, tensor-map descriptor and NCCL buffer
against its allocation table.  Nothing numerical happens here - that is what the GPU parity tests are for - but the
sequencing, the pointer arithmetic and the resource handling of every training entry point run for real, including the
opt-in paths that were written when no GPU was available ("Switches"): the launch sequences below are what the
first GPU run has to reproduce, and a descriptor or a slab pointer that leaves its buffer fails here first.

Each scenario runs tests/dryrun_driver.py in a process of its own (environment switches are read once per process).
The dry-run library is a test artefact under build/dryrun; the product only ever loads keras_unsupervised_b200/libkucd.so."""
import json
import os
import subprocess
import sys
from collections import Counter

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRY = os.path.join(ROOT, "build", "dryrun")
CSRC = os.path.join(ROOT, "keras_unsupervised_b200", "csrc")
FAKE_SRC = os.path.join(ROOT, "tools", "dryrun", "fake_cudart.cpp")


def _newer(target, sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


@pytest.fixture(scope="session")
def dry_build():
    """g++ the fake runtime, nvcc the engine against it (-cudart none): ~1 minute when the sources changed."""
    os.makedirs(DRY, exist_ok=True)
    fake = os.path.join(DRY, "libfakecudart.so")
    eng = os.path.join(DRY, "libkucd_dry.so")
    if not _newer(fake, [FAKE_SRC]):
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", fake, FAKE_SRC],
                       check=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "kucd.h"), fake]
    if not _newer(eng, srcs):
        subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "none",
                        "-shared", "-Xcompiler", "-fPIC", "-o", eng, os.path.join(CSRC, "kucd.cu"), "-L" + DRY,
                        "-lfakecudart", "-ldl", "-Xlinker", "-rpath=$ORIGIN"], check=True)
    return eng


def run(scenario, **env):
    e = {k: v for k, v in os.environ.items() if not k.startswith("KUCD_")}
    e.update({k: str(v) for k, v in env.items()})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_driver.py"), scenario], env=e,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def clean(snap):
    assert snap["errors"] == [], snap["errors"]
    return snap


def test_every_contraction_launch_is_decoded_and_checked(dry_build):
    """The fake reads GemmParams / ChainParams (csrc/params.h) out of each launch and checks what the epilogue may touch:
    bias (a whole 32-column chunk is read), state / probability planes, raw tiles or owner slots, column / row sums,
    injected draws, the chain's completion counters.  Live: the counter moves - and a block with nothing behind its
    pointers is refused."""
    d = run("cd_step")
    assert clean(d["bf16"])["decoded"] == 1 and clean(d["f32"])["decoded"] == 2    # one chain launch; at float32 grade the chain + dW
    import ctypes as C

    fake = C.CDLL(os.path.join(DRY, "libfakecudart.so"))
    fake.fake_reset()
    name = b"_ZN4kucd16gemm_bf16_kernelILi64ELb0ELb1ELi3ELi0ELi1EEEvNS_10GemmParamsE"      # free-energy epilogue
    stub = C.create_string_buffer(8)                                        # any unique address stands for the host stub
    fake.__cudaRegisterFunction(None, stub, None, name, -1, None, None, None, None, None)
    blob = (C.c_char * 4096)()
    C.memmove(C.addressof(blob) + 2560, (C.c_int32 * 6)(1, 0, 64, 64, 1, 0), 24)     # num_seg, neg_mask, M, N, kblocks, pad
    args = (C.c_void_p * 1)(C.addressof(blob))

    class Dim3(C.Structure):
        _fields_ = [("x", C.c_uint), ("y", C.c_uint), ("z", C.c_uint)]

    fake.cudaLaunchKernel.argtypes = [C.c_void_p, Dim3, Dim3, C.c_void_p, C.c_size_t, C.c_void_p]
    assert fake.cudaLaunchKernel(C.addressof(stub), Dim3(1, 1, 1), Dim3(320, 1, 1), args, 1024, None) == 0
    # the two operand descriptors were never encoded; bias and rowsum point nowhere
    assert fake.fake_counter(5) == 1 and fake.fake_error_count() == 4


CHAIN_SMALL = "chain_kernel<64,1,0,0>"
DW = "gemm_bf16_kernel<64,1,1,0,0,1>"          # A MN-major, B MN-major, raw epilogue: the dW contraction
DW16 = "gemm_bf16_kernel<128,1,1,6,0,1>"       # ... with the bf16 push epilogue (BN >= 128)
PROJ = ["gemm_bf16_kernel<64,0,1,1,0,1>", "gemm_bf16_kernel<64,0,0,1,0,1>", "gemm_bf16_kernel<64,0,1,2,0,1>"]  # h, v, h prob


def test_default_paths_launch_what_design_md_says(dry_build):
    d = run("cd_step")
    assert clean(d["bf16"])["kernels"] == ["ingest_kernel", "colsum_kernel", CHAIN_SMALL, "update_w_kernel<0>"]
    f32 = clean(d["f32"])["kernels"]                     # float32-grade: the chain kernel with piecewise accumulation
    assert f32 == ["ingest_kernel", "colsum_kernel", "chain_kernel<64,1,0,8>", "gemm_bf16_kernel<128,1,1,0,8,1>",
                   "update_w_kernel<0>"]                 # (CH = 8), then the four-term dW contraction on its own
    e = clean(run("fit_epoch"))
    assert e["graph"] == ["colsum_store_kernel", "memset", CHAIN_SMALL, "update_w_kernel<0>"]   # 3 kernels per replay
    assert e["kernels"] == ["set_dyn_kernel"] + ["graph_launch"] * 8 and e["steps"] == 8
    assert e["timings"]["graph_kernel_launches"] == 24


def test_two_chain_split_forks_and_joins_inside_the_capture(dry_build):
    d = clean(run("split", KUCD_SPLIT=2, KUCD_CHAIN=0))      # no capture-isolation / unjoined-work complaint from the fake
    g = d["graph"]
    assert d["kernels"] == ["set_dyn_kernel", "graph_launch", "graph_launch"]          # nothing escaped the capture
    assert len([k for k in g if k.startswith("gemm_bf16_kernel<64,0,")]) == 10        # 2 chains x (2k + 1) projections
    assert g[-2:] == [DW, "update_w_kernel<0>"]
    streams = {line.split("stream=")[1].split()[0] for line in d["graph_raw"] if line.startswith("launch")}
    assert len(streams) == 2                                                           # both streams are in the graph


def test_streamed_fit_per_minibatch_and_chunked(dry_build):
    plain = clean(run("fit_host", KUCD_STREAM_CHUNK=0))                 # per-minibatch stream (chunks of 8 are the default)
    c = Counter(plain["kernels"])
    assert c["ingest_kernel"] == 8 and c[CHAIN_SMALL] == 8 and c["update_w_kernel<0>"] == 8 and "graph_launch" not in c
    assert plain["timings"]["h2d_bytes"] == 1000 * 300 * 4 and plain["timings"]["d2h_bytes"] == 8 * 4
    assert clean(plain["second"])["mallocs"] == 0                       # a second pass allocates nothing
    chunked = clean(run("fit_host", KUCD_STREAM_CHUNK=4))               # 1000 rows, batch 128: chunks of 512 and 488 rows
    c = Counter(chunked["kernels"])
    assert c["ingest_kernel"] == 2 and c["set_dyn_kernel"] == 2 and c["graph_launch"] == 7
    assert c[CHAIN_SMALL] == 1 and c["update_w_kernel<0>"] == 1 and c["recon_small_kernel"] == 1   # the 104-row remainder
    # (statistic + its write into the page-locked log: one launch of one block)
    assert chunked["graph"] == ["colsum_store_kernel", "memset", CHAIN_SMALL, "recon_small_kernel", "update_w_kernel<0>"]
    assert chunked["timings"]["h2d_bytes"] == plain["timings"]["h2d_bytes"]
    assert chunked["timings"]["d2h_bytes"] == plain["timings"]["d2h_bytes"] and chunked["steps"] == 8
    second = clean(chunked["second"])
    assert second["mallocs"] == 0 and Counter(second["kernels"])["graph_launch"] == 7          # the captured step is kept


def test_data_set_planes_come_from_the_stream_ordered_pool(dry_build):
    """transform_dataset / inv_transform_dataset in a loop: the output planes are cudaMallocAsync'ed and freed in stream
    order (no cudaMalloc, no synchronising cudaFree per call), and nothing is left at context close."""
    d = clean(run("transform_loop"))
    assert d["mallocs"] == 0 and d["frees"] == 0 and d["async_allocs"] == 20 and d["live_after_close"] == 0


def test_delta_rule_is_projection_colsum_dw_update(dry_build):
    d = run("delta_rule")
    for key, proj in (("bf16_fwd", "gemm_bf16_kernel<64,0,1,2,0,1>"), ("bf16_bwd", "gemm_bf16_kernel<64,0,0,2,0,1>"),
                      ("f32_fwd", "gemm_bf16_kernel<128,0,1,2,8,1>"), ("f32_bwd", "gemm_bf16_kernel<128,0,0,2,8,1>")):
        k = clean(d[key])["kernels"]
        assert k[:2] == ["ingest_kernel", "ingest_kernel"] and k[2] == proj and k[3] == "colsum_kernel"
        assert k[4].startswith("gemm_bf16_kernel<") and ",1,1,0," in k[4] and k[5] == "update_w_kernel<0>" and len(k) == 6


def test_slabs_need_a_collective(dry_build):
    """KUCD_AR_SLABS pipelines the all-reduce of dW; on a single rank there is nothing to pipeline (the variant that hid the
    update behind the next slab's contraction was measured and removed): the switch is ignored."""
    off = clean(run("slabs"))
    assert off["graph"] == ["colsum_store_kernel", "memset", CHAIN_SMALL, "update_w_kernel<0>"]
    on = clean(run("slabs", KUCD_AR_SLABS=3, KUCD_AR_SLABS_MIN_ELEMS=1))
    assert on["graph"] == off["graph"]


@pytest.mark.parametrize("env,dw,exchange,update", [
    (dict(KUCD_FUSED_MIN_ROWS=1), [DW],
     ["push_bias_kernel", "peer_barrier_kernel", "reduce_bias_kernel"], ["update_w_sharded_kernel<0>"]),
    (dict(KUCD_FUSED_MIN_ROWS=1, KUCD_WIRE_BF16=1), [DW16],
     ["push_bias_kernel", "peer_barrier_kernel", "reduce_bias_kernel"], ["update_w_sharded_kernel<1>"]),
    (dict(KUCD_FUSED_REDUCE=0, KUCD_WIRE_BF16=1), [DW16], ["allreduce", "allreduce"], ["update_w_kernel<1>"]),
    (dict(KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=2, KUCD_AR_SLABS_MIN_ELEMS=1), [DW, "allreduce", "allreduce", DW], ["allreduce"],
     ["update_w_kernel<0>", "update_w_kernel<0>"]),
    (dict(KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=2, KUCD_AR_SLABS_MIN_ELEMS=1, KUCD_WIRE_BF16=1),
     [DW16, "allreduce", "allreduce", DW16], ["allreduce"], ["update_w_kernel<1>", "update_w_kernel<1>"]),
])
def test_two_rank_exchange_variants(dry_build, env, dw, exchange, update):
    d = clean(run("two_ranks", **env))
    g = d["graph"]
    assert g[:4] == ["colsum_store_kernel"] + PROJ
    assert g[4:4 + len(dw) + len(exchange) + len(update)] == dw + exchange + update
    for t in d["timings"]:
        assert t["graph_launches"] == 4
    if "allreduce" in exchange + dw:   # the counters count executed steps (4 replays), not the one capture
        assert d["timings"][0]["allreduce_calls"] == 4 and d["timings"][0]["fused_reduce_steps"] == 0
    else:
        assert d["timings"][0]["fused_reduce_steps"] == 4 and d["timings"][0]["allreduce_calls"] == 0


def test_checkpoint_state_through_the_abi(dry_build):
    """kucd_rbm_get/set_momentum (pitched copies of the padded fp32 buffers; None until a momentum step or a set has created
    them), kucd_rbm_get/set_draw_counters (reset by set_seed), shape errors as ValueError - the fake runtime really copies."""
    d = run("momentum")
    assert d["absent"] is True and d["round_trip"] is True and d["used"] is True and d["created_by_step"] is True
    assert d["draws"] == {"infer_draws": 7, "score_draws": 9}
    assert d["draws_after_set_seed"] == {"infer_draws": 0, "score_draws": 0}
    assert "momentum of rbm_weight has shape (333, 100), expected (333, 130)" in d["bad_shape"]
    clean(d["snapshot"])


def test_which_exchange_the_default_rule_picks(dry_build):
    """choose_exchange without size overrides.  Between the two data-parallel exchanges (KUCD_EXCHANGE=dp), C3's weight matrix
    (4096 x 4096): shards of <= 256 rows keep the all-reduce, shards of 257 ... 2047 rows take the fused exchange (their
    projections run as a mid chain; measured at C3 strong-scaled over 8 GPUs: 0.833 vs 1.002 ms per step), as do shards of
    >= 2048 rows.  A small model (784 x 500) keeps the all-reduce below 2048 rows: its step is one small-tile chain launch with
    dW inside.  With nothing forced, a CD-1 step on the large matrix goes unit-sharded (4 bit exchanges against 64 MiB of dW)."""
    for V, H, rows, fused in ((4096, 4096, 256, False), (4096, 4096, 512, True), (4096, 4096, 1024, True),
                              (784, 500, 64, False), (784, 500, 512, False), (784, 500, 1024, False), (784, 500, 2048, True)):
        t = clean(run("two_ranks", DRY_ROWS=rows, DRY_V=V, DRY_H=H, KUCD_EXCHANGE="dp"))["timings"][0]
        assert t["graph_launches"] == 4 and t["unit_steps"] == 0
        assert (t["fused_reduce_steps"], t["allreduce_calls"]) == ((4, 0) if fused else (0, 4)), (V, H, rows, t)
    t = clean(run("two_ranks", DRY_ROWS=512, DRY_V=4096, DRY_H=4096))["timings"][0]
    assert (t["unit_steps"], t["fused_reduce_steps"], t["allreduce_calls"]) == (4, 0, 0)


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_unit_sharded_step_over_in_process_ranks(dry_build, ranks):
    """KUCD_EXCHANGE=units (kucd.cu: enqueue_cd_units): per projection one launch over the rank's slice of the units and all
    rows of the global minibatch, then pack + peer stores, flag barrier, expansion; dW over the owned columns, the
    unit-sharded update, a closing barrier.  Every tensor map, slice pointer, bias chunk and column-sum pointer of the
    sliced launches passes the fake's checks; a range with a remainder minibatch falls back to the all-reduce."""
    d = run("units", KUCD_EXCHANGE="units", DRY_RANKS=ranks)
    for part in (d, d["pcd"], d["remainder"]):
        assert part["errors"] == [], part["errors"][:10]
    ex = ["pack_push_kernel", "peer_barrier_kernel", "ingest_bits_kernel"]
    g = [k.split("<")[0] for k in d["graph"]]
    gemm = "gemm_bf16_kernel"
    assert g == ["memset"] + ex + ["colsum_kernel", gemm] + ex + [gemm] + ex + [gemm] + ex + [gemm] + ex + [gemm, gemm] + \
        ["update_w_units_kernel", "update_bias_kernel", "update_bias_kernel", "peer_barrier_kernel", "advance_dyn_kernel"]
    gp = [k.split("<")[0] for k in d["pcd"]["graph"]]
    assert gp.count("pack_push_kernel") == 3 and gp.count(gemm) == 5    # h_pos is not exchanged under PCD and "copy_rows_kernel" not in gp
    assert "allreduce" in d["remainder"]["graph"] and "pack_push_kernel" not in d["remainder"]["graph"]
    t = d["timings"][0]
    assert t["unit_steps"] == 3 + 3 + 1 and t["unit_exchanges"] == 3 * 5 + 3 * 3 + 3 + 2   # + one gather of the chains per training call


def _allreduces(d):
    """(count, nccl dtype) of every all-reduce of the captured step"""
    out = []
    for line in d["graph_raw"]:
        if line.startswith("allreduce"):
            f = dict(x.split("=") for x in line.split()[1:])
            out.append((int(f["count"]), int(f["dtype"])))
    return out


def test_what_the_all_reduces_carry(dry_build):
    V, H, ldH = 784, 500, 512
    bias = (1024 + 256) + (512 + 256)                                   # [db | dc], each padded to 256 + 256 (kucd.cu)
    F32, BF16 = 7, 9                                                    # ncclFloat32, ncclBfloat16
    d = clean(run("two_ranks", KUCD_FUSED_REDUCE=0))
    assert d["graph"] == ["colsum_store_kernel", "memset", CHAIN_SMALL, "allreduce", "update_w_kernel<0>"]
    assert _allreduces(d) == [(V * ldH + bias, F32)]                    # one call over the whole block [dW | db | dc]
    d = clean(run("two_ranks", KUCD_FUSED_REDUCE=0, KUCD_WIRE_BF16=1))
    assert _allreduces(d) == [(V * ldH, BF16), (bias, F32)]             # dW as bf16, the small statistics as float32
    d = clean(run("two_ranks", KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=2, KUCD_AR_SLABS_MIN_ELEMS=1, KUCD_WIRE_BF16=1))
    assert _allreduces(d) == [(512 * ldH, BF16), (bias, F32), ((V - 512) * ldH, BF16)]   # slabs of 512 and 272 rows
    d = clean(run("two_ranks", KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=2, KUCD_AR_SLABS_MIN_ELEMS=1))
    assert _allreduces(d) == [(512 * ldH, F32), (bias, F32), ((V - 512) * ldH, F32)]


def test_the_fake_runtime_does_catch_a_bad_descriptor(dry_build):
    """The checker checks: a tensor map whose extent leaves its allocation is refused."""
    import ctypes as C

    fake = C.CDLL(os.path.join(DRY, "libfakecudart.so"))
    fake.fake_reset()
    p = C.c_void_p()
    assert fake.cudaMalloc(C.byref(p), C.c_size_t(64 * 128 * 2)) == 0
    fn = C.c_void_p()
    res = C.c_int()
    assert fake.cudaGetDriverEntryPoint(b"cuTensorMapEncodeTiled", C.byref(fn), C.c_ulonglong(0), C.byref(res)) == 0
    enc = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_int, C.c_int)(fn.value)
    tm = C.create_string_buffer(128 + 64)
    tm_aligned = (C.addressof(tm) + 63) & ~63
    ones = (C.c_uint32 * 2)(1, 1)
    box = (C.c_uint32 * 2)(64, 64)
    strides = (C.c_uint64 * 1)(128 * 2)
    BF16, SW128 = 9, 3                                                  # CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, SWIZZLE_128B
    good = (C.c_uint64 * 2)(128, 64)
    assert enc(tm_aligned, BF16, 2, p, good, strides, box, ones, 0, SW128, 2, 0) == 0
    bad_rows = (C.c_uint64 * 2)(128, 65)                                # one row more than was allocated
    assert enc(tm_aligned, BF16, 2, p, bad_rows, strides, box, ones, 0, SW128, 2, 0) != 0
    assert enc(tm_aligned, BF16, 2, C.c_void_p(p.value + 8), good, strides, box, ones, 0, SW128, 2, 0) != 0   # misaligned
    assert fake.fake_error_count() == 2
    assert fake.cudaFree(p) == 0


def test_the_fake_runtime_does_catch_capture_mistakes(dry_build):
    """Unjoined forked work, waiting on an event from outside the capture, and waiting on a never-recorded event."""
    import ctypes as C

    fake = C.CDLL(os.path.join(DRY, "libfakecudart.so"))
    fake.fake_reset()
    a, b = C.c_void_p(), C.c_void_p()
    e_in, e_out, e_never, g = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    p = C.c_void_p()
    for s in (a, b):
        assert fake.cudaStreamCreateWithFlags(C.byref(s), 1) == 0
    for e in (e_in, e_out, e_never):
        assert fake.cudaEventCreate(C.byref(e)) == 0
    assert fake.cudaMalloc(C.byref(p), C.c_size_t(4096)) == 0
    assert fake.cudaEventRecord(e_out, b) == 0                               # recorded before the capture
    assert fake.cudaStreamBeginCapture(a, 2) == 0
    assert fake.cudaStreamWaitEvent(a, e_out, 0) != 0                        # isolation
    assert fake.cudaStreamWaitEvent(a, e_never, 0) == 0                      # legal, but orders nothing: reported
    assert fake.cudaEventRecord(e_in, a) == 0
    assert fake.cudaStreamWaitEvent(b, e_in, 0) == 0                         # fork
    assert fake.cudaMemsetAsync(p, 0, C.c_size_t(4096), b) == 0              # captured work on the forked stream ...
    assert fake.cudaStreamEndCapture(a, C.byref(g)) == 0                     # ... never joined
    assert fake.fake_error_count() == 3
    msgs = []
    buf = C.create_string_buffer(512)
    for i in range(3):
        fake.fake_error_line(i, buf, 512)
        msgs.append(buf.value.decode())
    assert "Isolation" in msgs[0] and "never recorded" in msgs[1] and "Unjoined" in msgs[2]
    # the same with a join is clean
    fake.fake_reset()
    assert fake.cudaStreamBeginCapture(a, 2) == 0
    assert fake.cudaEventRecord(e_in, a) == 0
    assert fake.cudaStreamWaitEvent(b, e_in, 0) == 0
    assert fake.cudaMemsetAsync(p, 0, C.c_size_t(4096), b) == 0
    assert fake.cudaEventRecord(e_out, b) == 0
    assert fake.cudaStreamWaitEvent(a, e_out, 0) == 0
    assert fake.cudaStreamEndCapture(a, C.byref(g)) == 0
    assert fake.fake_error_count() == 0 and fake.fake_last_graph_size() == 5      # the memset + 2 records + 2 waits
    assert fake.cudaFree(p) == 0


def test_full_size_shapes(dry_build):
    """BASELINE.json's C3 (4096 -> 4096, batch 4096, CD-10) and C4 (PCD 16384 -> 8192, 1024 rows per GPU): every
    descriptor and pointer stays inside its buffer at full size, and the step is the whole-chain kernel on CTA pairs, the
    dW contraction on CTA pairs and the update."""
    d = run("full_size")
    c3, c4 = clean(d["c3"]), clean(d["c4"])
    chain, dw = "chain_kernel<256,2,0,0>", "gemm_bf16_kernel<256,1,1,0,0,2>"
    assert c3["graph"] == ["memset", "colsum_kernel", "memset", chain, dw, "update_w_kernel<0>", "advance_dyn_kernel"]
    assert c4["graph"] == ["memset", "colsum_kernel", "memset", chain, dw, "copy_rows_kernel", "update_w_kernel<0>",
                           "advance_dyn_kernel"]
    lines = [x for x in c3["graph_raw"] if "chain_kernel" in x]
    assert "grid=148,1,1" in lines[0] and "block=320,1,1" in lines[0] and "cluster=2" in lines[0]   # 74 CTA pairs


def test_reference_facing_classes_end_to_end(dry_build):
    """DBN.fit / transform / inv_transform / fine_tune / generate and a one-epoch float32 RBM.fit + free energy through the
    real ctypes layer and the real host code (values are meaningless: no kernel runs)."""
    d = run("python_surface", KUCD_STREAM_CHUNK=0)            # per-minibatch stream (chunking has its own test)
    for key in ("fit", "transform", "fine_tune", "generate", "one_epoch_f32"):
        clean(d[key])
    assert d["shapes"] == [[64, 2000], [64, 784]] and d["generate_shape"] == [16, 784] and d["fe_shape"] == [50]
    fit = Counter(d["fit"]["kernels"])
    assert fit["graph_launch"] == 3 * 2 * 5                 # 3 layers x 2 epochs x ceil(600 / 128) replayed steps
    assert fit["gemm_bf16_kernel<64,0,1,1,0,1>"] == 2       # the two inter-layer transforms of the whole data set (dbn.py:55)
    ft = Counter(d["fine_tune"]["kernels"])                 # 256 rows = 2 minibatches, 3 layers
    assert ft[CHAIN_SMALL] == 2                             # CD in the top RBM
    assert ft[DW] == 8 and ft["colsum_kernel"] == 8         # 2 lower layers x (generative + recognition) x 2 minibatches
    assert ft["gemm_bf16_kernel<64,0,1,2,0,1>"] == 4 and ft["gemm_bf16_kernel<64,0,0,2,0,1>"] == 4   # their predictions
    assert ft["update_w_kernel<0>"] == 10
    f32 = Counter(d["one_epoch_f32"]["kernels"])            # 1000 rows, batch 128: 8 streamed float32-grade steps
    assert f32["update_w_kernel<0>"] == 8 and f32["gemm_bf16_kernel<128,1,1,0,8,1>"] == 8


def test_data_formats_gaussian_pcd_injection_score(dry_build):
    """The formats either side of the path (packed bits in and out, uint8, the epoch shuffle into a new and into an existing
    data set), Gaussian visibles in both compute modes, persistent chains, injected draws, statistics and the score chain:
    every ingest / export / permute / copy launch has its source and destination extents checked by the fake."""
    d = run("data_formats", KUCD_STREAM_CHUNK=0)
    p = Counter(clean(d["packed"])["kernels"])
    assert d["shapes"] == [[500, 42], [500, 333]] and d["bits_out"] == [[500, 17], 130]
    assert p["ingest_bits_kernel"] == 4 + 1 + 1             # 4 streamed minibatches, the data set, the transform input
    assert p["ingest_kernel"] == 4 and p["permute_rows_kernel"] == 2 and p["export_bits_kernel"] == 2
    assert p[CHAIN_SMALL] == 8 and p["update_w_kernel<0>"] == 8
    g = Counter(clean(d["gaussian"])["kernels"])
    assert g["chain_kernel<64,1,1,0>"] == 1                   # the Gaussian instantiation of the small chain kernel
    assert g["gemm_bf16_kernel<128,0,1,4,8,1>"] == 1 and g["gemm_bf16_kernel<128,0,0,5,8,1>"] == 1   # relu / normal epilogues
    s = Counter(clean(d["pcd_inject_score"])["kernels"])
    assert s["copy_rows_kernel"] == 1 and s["score_kernel"] == 2 and s["free_energy_finish_kernel"] == 4


@pytest.mark.parametrize("env", [
    {}, {"KUCD_STREAM_CHUNK": 4}, {"KUCD_STREAM_CHUNK": 64}, {"KUCD_STREAM_GRAPH": 1},
    {"KUCD_AR_SLABS": 3, "KUCD_AR_SLABS_MIN_ELEMS": 1}, {"KUCD_PLANE_POOL": 1}, {"KUCD_CHAIN": 0}, {"KUCD_SMALL_CHAIN": 0},
    {"KUCD_CHAIN": 2}, {"KUCD_SPLIT": 2, "KUCD_CHAIN": 0}, {"KUCD_MERGE": 0}, {"KUCD_CG": 1}, {"KUCD_CHAIN_DW": 1},
], ids=lambda e: " ".join("%s=%s" % kv for kv in e.items()) or "defaults")
def test_sweep_of_shapes_and_options_under_every_switch(dry_build, env):
    """840 training calls - four model shapes (ragged and tile-sized), both compute modes, five data-set / minibatch size
    pairs (remainders, fewer rows than one minibatch), CD-1 / CD-3 with momentum and mean normalisation / PCD, host rows
    as float32, uint8, pitched and bit-packed, resident data sets, partial ranges, single steps - under each schedule
    switch: no copy, descriptor, epilogue buffer, data-path kernel, capture or event complaint from the fake."""
    d = run("sweep", **env)
    assert d["errors"] == [], d["errors"][:10]
    assert d["runs"] == 840 and d["decoded"] > 1500


@pytest.mark.parametrize("env,fused", [
    (dict(DRY_RANKS=2, KUCD_FUSED_MIN_ROWS=1), True), (dict(DRY_RANKS=2, KUCD_FUSED_MIN_ROWS=1, KUCD_WIRE_BF16=1), True),
    (dict(DRY_RANKS=8, KUCD_FUSED_MIN_ROWS=1), True), (dict(DRY_RANKS=8, KUCD_FUSED_MIN_ROWS=1, KUCD_WIRE_BF16=1), True),
    (dict(DRY_RANKS=3, KUCD_FUSED_MIN_ROWS=1, KUCD_WIRE_BF16=1), True),
    (dict(DRY_RANKS=2), False),                                     # 64 / 256 rows per rank: below the fused threshold
    (dict(DRY_RANKS=4, KUCD_FUSED_REDUCE=0), False), (dict(DRY_RANKS=4, KUCD_FUSED_REDUCE=0, KUCD_WIRE_BF16=1), False),
    (dict(DRY_RANKS=4, KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=3, KUCD_AR_SLABS_MIN_ELEMS=1), False),
    (dict(DRY_RANKS=4, KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=3, KUCD_AR_SLABS_MIN_ELEMS=1, KUCD_WIRE_BF16=1), False),
], ids=lambda e: " ".join("%s=%s" % kv for kv in e.items()) if isinstance(e, dict) else str(e))
def test_sweep_of_the_exchange_variants_over_in_process_ranks(dry_build, env, fused):
    """2, 3, 4 and 8 contexts of one process as the ranks of a group (the fake NCCL talks to nobody, the fake IPC passes
    pointers through, so the owners' slots of the fused exchange really are each other's memory): four model shapes -
    visible counts that do and do not divide by the rank count - resident fits with a remainder minibatch, streamed fits,
    single steps, CD-1 and CD-2 with momentum and mean normalisation; fused exchange with fp32 and bf16 slots, NCCL with
    fp32 and bf16 payloads, slab-pipelined.  Every slot pointer, slab pointer and all-reduce buffer stays inside its
    allocation."""
    d = run("ranks_sweep", **env)
    assert d["errors"] == [], d["errors"][:10]
    assert d["runs"] == 48 * int(env["DRY_RANKS"])
    t = d["timings"]
    assert (t["fused_reduce_steps"] > 0 and t["allreduce_calls"] == 0) if fused else \
        (t["fused_reduce_steps"] == 0 and t["allreduce_calls"] > 0)


def test_the_fake_runtime_does_catch_a_replay_into_freed_memory(dry_build):
    import ctypes as C

    fake = C.CDLL(os.path.join(DRY, "libfakecudart.so"))
    fake.fake_reset()
    a, g, ge, p = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert fake.cudaStreamCreateWithFlags(C.byref(a), 1) == 0
    assert fake.cudaMalloc(C.byref(p), C.c_size_t(4096)) == 0
    assert fake.cudaStreamBeginCapture(a, 2) == 0
    assert fake.cudaMemsetAsync(p, 0, C.c_size_t(4096), a) == 0
    assert fake.cudaStreamEndCapture(a, C.byref(g)) == 0
    assert fake.cudaGraphInstantiate(C.byref(ge), g, C.c_ulonglong(0)) == 0
    assert fake.cudaGraphLaunch(ge, a) == 0 and fake.fake_error_count() == 0
    assert fake.cudaFree(p) == 0
    assert fake.cudaGraphLaunch(ge, a) == 0 and fake.fake_error_count() == 1       # the buffer behind the graph is gone


def test_argument_errors_of_the_c_abi(dry_build):
    """Shape, range and state errors come back as KUCD_ERR_INVALID_ARG / SHAPE_MISMATCH -> ValueError with the message of
    kucd_last_error (the reference raises ValueError for its own argument errors, dbn.py:29,48), before anything is
    launched, and nothing stays allocated behind a failed call."""
    d = clean(run("errors"))
    seen = d["seen"]
    accepted = [i for i, (kind, _) in enumerate(seen) if kind == "no error"]
    assert accepted == [11]                                  # an integer array: the ctypes shim converts it to float32
    assert all(kind == "ValueError" for i, (kind, _) in enumerate(seen) if i != 11), seen
    msgs = [m for _, m in seen]
    assert "expected (-1, 64)" in msgs[0] and "k = 0 is outside" in msgs[4] and "chains" in msgs[6]
    assert "shuffled in place" in msgs[15] and "MODE_COMPLEX" in msgs[17] and msgs[18] == "device 7 of 2"
    assert d["leaked"] == 0
    assert d["kernels"] == ["ingest_kernel", "colsum_store_kernel", "chain_kernel<64,1,0,0>", "update_w_kernel<0>",
                            "refresh_planes_kernel"]          # the accepted call, and the second model's set_params


def test_allocation_failures_anywhere_in_a_session(dry_build):
    """Fault injection: the n-th cudaMalloc of a session (context, model, data set, resident fit with momentum, data-set
    transform, shuffle, streamed fit, delta rule, transform, free energy; both compute modes) fails, for n = 1 .. 45.  The
    call that hits it raises KucdError (KUCD_ERR_CUDA), the process survives, closing the context gives everything back."""
    d = run("oom")
    outcomes = Counter(r[1] for r in d["results"])
    assert set(outcomes) <= {"KucdError", "ok"} and outcomes["KucdError"] >= 35, d["results"]
    assert all(leak == 0 and complaints == 0 for _, _, leak, complaints in d["results"]), d["results"]


def _order(graph_raw):
    """The captured step as (what, stream, event) triples with streams / events renamed s0, s1 .. / e0, e1 .. by appearance."""
    import re

    names = {}

    def nm(h, prefix):
        return names.setdefault((prefix, h), "%s%d" % (prefix, sum(1 for k in names if k[0] == prefix)))

    out = []
    for line in graph_raw:
        st = re.search(r"stream=(\S+)", line)
        ev = re.search(r"event=(\S+)", line)
        what = line.split()[0]
        if what == "launch":
            what = re.sub(r"^_ZN4kucd\d+", "", line.split()[1])
            what = re.match(r"[a-z_0-9]+", what).group(0)
        out.append((what, nm(st.group(1), "s") if st else None, nm(ev.group(1), "e") if ev else None))
    return out


def test_slab_ordering_inside_the_captured_step(dry_build):
    """Two ranks, NCCL, two slabs: how the slab-pipelined all-reduce is ordered inside the captured step."""
    # two ranks, NCCL: the all-reduce of slab i sits on the second stream behind contraction i, update i behind all-reduce i
    d = clean(run("two_ranks", KUCD_FUSED_REDUCE=0, KUCD_AR_SLABS=2, KUCD_AR_SLABS_MIN_ELEMS=1))
    o = [x for x in _order(d["graph_raw"]) if x[0] in ("gemm_bf16_kernel", "allreduce", "update_w_kernel", "event_record", "event_wait")]
    o = o[3:]                                                # the three projections
    assert o == [("gemm_bf16_kernel", "s0", None), ("event_record", "s0", "e0"), ("event_wait", "s1", "e0"),
                 ("allreduce", "s1", None), ("allreduce", "s1", None), ("event_record", "s1", "e1"),
                 ("gemm_bf16_kernel", "s0", None), ("event_record", "s0", "e2"), ("event_wait", "s1", "e2"),
                 ("allreduce", "s1", None), ("event_record", "s1", "e3"),
                 ("event_wait", "s0", "e1"), ("update_w_kernel", "s0", None),
                 ("event_wait", "s0", "e3"), ("update_w_kernel", "s0", None)]


def test_rbm_fit_under_the_optional_hps_keys(dry_build):
    """RBM.fit through the real ctypes layer and host code: shuffling epochs, the reference's three single-parameter runs +
    score chain per minibatch (rbm.py:214-234), persistent chains with momentum / weight decay / mean normalisation, the
    constructor-default Gaussian mode in float32-grade arithmetic, and a one-epoch fit of an uploaded array."""
    d = run("rbm_options", KUCD_STREAM_CHUNK=0)
    sh = Counter(clean(d["shuffle"])["kernels"])
    assert sh["permute_rows_kernel"] == 3 and sh["graph_launch"] == 3 * 5 and d["shuffle"]["history"] == 3
    ref = Counter(clean(d["reference"])["kernels"])          # 600 rows / 128 = 5 minibatches
    assert ref["update_w_kernel<0>"] == 15                   # runs A, B, C: one parameter each (rbm.py:214-216)
    assert ref["gemm_bf16_kernel<128,1,1,0,8,1>"] == 15      # each with its own chain and statistics
    assert ref["score_kernel"] == 5 and ref["free_energy_finish_kernel"] == 10     # run D: F(v), F(v_neg) (rbm.py:227-233)
    pcd = Counter(clean(d["pcd"])["kernels"])
    assert pcd["graph_launch"] == 10 and pcd["ingest_kernel"] == 2                 # the data set and the chains' start
    g = Counter(clean(d["gaussian_default"])["kernels"])
    assert g["gemm_bf16_kernel<128,0,1,4,8,1>"] == 5 and g["gemm_bf16_kernel<128,0,0,5,8,1>"] == 5
    one = Counter(clean(d["resident_one_epoch"])["kernels"])
    assert one["graph_launch"] == 5 and one["ingest_kernel"] == 1
