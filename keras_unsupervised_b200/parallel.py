"""Host-side logic of data-parallel training (one process per GPU).

The minibatch shards by rows: given W, b, c every row's Gibbs chain is independent (rbm.py:119-124) and
rows only meet in the batch sums dW, dc, db (rbm.py:125-134), which the engine all-reduces over NCCL.
Rank r owns rows [r*b, (r+1)*b) of every global minibatch, b = batch_size / world; Philox draws are
keyed by the row's index inside the GLOBAL minibatch, so n ranks sample exactly what one rank would.
"""
from __future__ import annotations


def local_batch(batch_size: int, n_rows: int, world: int) -> int:
    if world < 1:
        raise ValueError("world must be >= 1")
    if batch_size % world or (n_rows % batch_size) % world:
        raise ValueError(f"data-parallel fit over {world} ranks needs batch_size and the remainder minibatch "
                         f"to divide by {world} (batch_size={batch_size}, rows={n_rows})")
    return batch_size // world


def shard_rows(V, batch_size: int, rank: int, world: int):
    """-> (local rows laid out so that local minibatch i is this rank's slice of global minibatch i,
           local batch size, global_row0 = index of the rank's first row inside a global minibatch)."""
    if world == 1:
        return V, batch_size, 0
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    n_rows, dim = V.shape
    b = local_batch(batch_size, n_rows, world)
    full = n_rows // batch_size
    if type(V).__name__ == "PackedBits":  # rows are byte strings: shard the byte matrix, keep the column count
        data, _, row0 = shard_rows(V.data, batch_size, rank, world)
        return type(V)(data, V.n_cols), b, row0
    parts = []
    if full:
        parts.append(V[:full * batch_size].reshape(full, world, b, dim)[:, rank].reshape(full * b, dim))
    rem = n_rows - full * batch_size
    if rem:
        rb = rem // world
        parts.append(V[full * batch_size + rank * rb: full * batch_size + (rank + 1) * rb])
    if len(parts) == 1:
        local = parts[0]
    elif type(V).__module__.split(".")[0] == "torch":
        import torch

        local = torch.cat(parts)
    else:
        import numpy as np

        local = np.concatenate(parts)
    return local, b, rank * b
