"""Sizes for an ncu launch list of the data-path kernels (ingest_kernel<T>, ingest_bits_kernel, export_bits_kernel,
permute_rows_kernel) and, without ncu, the effect of the Gaussian-visible chain kernels.

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ingest|export_bits|permute' --csv \
        --log-file gpurun_out/launches_data_path.csv python tools/data_path_bench.py kernels
    python tools/data_path_bench.py gauss          # ms per CD-10 step, Gaussian visibles, chain kernel on / off

Algorithmic bytes per unit: ingest f32 4 + 2, u8 1 + 2, bits 1/8 + 2; export bits 2 + 1/8;
shuffle 2 + 2 (per live plane).  The sources are resident on the GPU, so the launch times are HBM times.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from keras_unsupervised_b200 import _lib as L  # noqa: E402
from keras_unsupervised_b200.data import PackedBits  # noqa: E402
from keras_unsupervised_b200.engine import Context, Dataset, Machine  # noqa: E402


def kernels():
    ctx = Context(device=0, seed=1)
    N, V = 262144, 4096
    gen = torch.Generator(device="cuda").manual_seed(3)
    bits = torch.randint(0, 256, (N, V // 8), device="cuda", dtype=torch.uint8, generator=gen)
    ds = Dataset.from_array(ctx, PackedBits(bits, V), L.COMPUTE_BF16)             # ingest_bits_kernel
    order = ds.shuffled(1, 0)                                                      # permute_rows_kernel
    out = PackedBits(torch.empty((N, V // 8), dtype=torch.uint8, device="cuda"), V)
    keep = []
    import ctypes as C
    L.check(ctx.lib.kucd_dataset_read(order.handle, C.byref(L.tensor_of(out, keep))))   # export_bits_kernel
    # a permutation of the rows moves every byte to another row of the same column: the column sums are unchanged
    # (that it is THE permutation the oracle defines is what tests/test_gpu_data_path.py checks)
    perm_ok = bool(torch.equal(out.data.to(torch.int64).sum(dim=0), bits.to(torch.int64).sum(dim=0)))
    order.close()
    ds.close()
    u8 = (torch.rand((N, V), device="cuda", generator=gen) < 0.5).to(torch.uint8)
    d8 = Dataset.from_array(ctx, u8, L.COMPUTE_BF16)                               # ingest_kernel<uint8_t>
    d8.close()
    f32 = u8[: N // 2].to(torch.float32)
    del u8
    d32 = Dataset.from_array(ctx, f32, L.COMPUTE_BF16)                             # ingest_kernel<float>, N/2 rows
    d32.close()
    print(json.dumps({"rows": N, "cols": V, "f32_rows": N // 2, "shuffle_keeps_column_sums": perm_ok}))


def gauss():
    res = {}
    for chain in ("1", "0"):
        os.environ["KUCD_CHAIN"] = chain
        ctx = Context(device=0, seed=1)
        for name, (V, H, B, k, steps) in {"c3_shape": (4096, 4096, 4096, 10, 30), "c1_shape": (784, 500, 128, 1, 400)}.items():
            m = Machine(ctx, V, H, L.MODE_VISIBLE_GAUSSIAN, L.COMPUTE_BF16, seed=5)
            rng = np.random.default_rng(0)
            m.set_params(rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
            X = torch.randn((8 * B, V), device="cuda")
            ds = Dataset.from_array(ctx, X, L.COMPUTE_BF16)
            hp = Machine.hparams(lr=1e-5, k=k, normalize=True)
            m.fit_range(ds, B, hp, 0, 8)
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(steps // 8):
                m.fit_range(ds, B, hp, 0, 8)
            ctx.sync()
            ms = 1e3 * (time.perf_counter() - t0) / (steps // 8 * 8)
            W = m.get_params()[0]
            res[f"{name}_chain{chain}"] = {"ms_per_step": round(ms, 4), "finite": bool(np.isfinite(W).all()),
                                           "chain_launches": ctx.timings()["chain_launches"]}
            ds.close()
            m.close()
        ctx.close()
    print(json.dumps(res))


if __name__ == "__main__":
    {"kernels": kernels, "gauss": gauss}[sys.argv[1]]()
