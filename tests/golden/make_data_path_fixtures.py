"""Freezes the two data formats that sit on disk / in checkpoints across runs: the packed-bit layout and the
shuffling stream (oracle/cd_oracle.py: pack_bits, feistel_keys, feistel_permutation).  Both are definitions of this
engine, not of the reference (which has neither: rbm.py:218 never shuffles, its data are float32), so the fixture
pins them against accidental change between rounds; tests/test_data_path.py compares the oracle - and, through
tools/data_path_host.cu, the device functions - with it.

    python tests/golden/make_data_path_fixtures.py     # rewrites tests/golden/data_path.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cd_oracle as O  # noqa: E402


def build():
    rng = np.random.default_rng(2026)
    x = (rng.random((3, 19)) < 0.5).astype(np.uint8)
    cases = [(10, 42, 0), (1000, 42, 7), (60000, 2**40 + 5, 2**33 + 1)]
    return {
        "pack": {"dense": x.tolist(), "packed": O.pack_bits(x).tolist()},
        "keys": [{"seed": s, "epoch": e, "keys": O.feistel_keys(s, e)} for _, s, e in cases],
        "perm": [{"rows": n, "seed": s, "epoch": e, "head": O.feistel_permutation(n, s, e)[:16].tolist(),
                  "checksum": int((O.feistel_permutation(n, s, e) * (np.arange(n) + 1)).sum() % (2**61 - 1))}
                 for n, s, e in cases],
    }


if __name__ == "__main__":
    with open(os.path.join(HERE, "data_path.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote", os.path.join(HERE, "data_path.json"))
