"""Deadlock-freedom of the whole-chain kernel's schedule (keras_unsupervised_b200/csrc/chain.cuh), checked on a model.

The kernel flattens every contraction of a CD-k minibatch into one tile sequence (stage-major, row-block-major inside a
stage); persistent workers (CTAs or CTA pairs) take tiles q = w, w + n, w + 2n, ... in order, and the producer of a
worker waits - before loading a tile's operands - until the row block(s) that tile reads have been completely written,
by whichever workers own those tiles.  The claim in the kernel's header: every wait is on tiles EARLIER in the sequence,
so the smallest unfinished tile can always run, for any number of resident workers.

`chain_plan` restates the stage table that `launch_chain` (csrc/kucd.cu) builds - stage kinds, dependencies, the dW
contraction as an optional last stage with its two K-segments - and `run` plays the workers: a worker is stuck on its
current tile until the tile's dependencies are complete; the schedule is sound if all tiles complete for every shape,
chain length and worker count.  A restatement, not the C++ itself: it guards the design (and whoever adds a stage kind
next - the float32-grade chain is on the list) rather than the build."""
import itertools

import pytest

KBLOCK_M = 128


def chain_plan(batch, V, H, k, pcd=False, with_dw=False, small=False):
    """-> list of stages: dict(num_m, num_n, dep, all_blocks, dep2).  dep: stage whose row block (same index) must be
    complete before the tile's operands are loaded, or ALL its row blocks when all_blocks (the dW contraction reads whole
    state matrices); dep2: stage that must be complete in all its row blocks before the second K-segment is loaded."""
    bn, cg = (64, 1) if small else (256, 2)
    tile_m = KBLOCK_M * cg
    num_m_batch = -(-batch // tile_m)
    fwd = dict(num_m=num_m_batch, num_n=-(-H // bn), all_blocks=False, dep2=None)
    bwd = dict(num_m=num_m_batch, num_n=-(-V // bn), all_blocks=False, dep2=None)
    stages = [dict(fwd, dep=None)]                    # h_pos from v0
    hsrc = 0
    if pcd:
        stages.append(dict(fwd, dep=None))            # first h of the stored chains
        hsrc = 1
    for _ in range(k):
        stages.append(dict(bwd, dep=hsrc))            # v from the current h
        stages.append(dict(fwd, dep=len(stages) - 1)) # next h (probability on the last step) from that v
        hsrc = len(stages) - 1
    if with_dw:
        stages.append(dict(num_m=-(-V // tile_m), num_n=-(-H // bn), dep=0, all_blocks=True, dep2=hsrc))
    return stages


def run(stages, workers):
    """Play the persistent workers; returns the number of tiles completed (== total iff no deadlock)."""
    tiles = []                                        # (stage, row block) of every tile, in sequence order
    for s, st in enumerate(stages):
        for m in range(st["num_m"]):
            tiles.extend((s, m) for _ in range(st["num_n"]))
    total = len(tiles)
    done = [[0] * st["num_m"] for st in stages]       # tiles finished per (stage, row block)

    def block_complete(s, m):
        return done[s][m] == stages[s]["num_n"]

    def stage_complete(s):
        return all(block_complete(s, m) for m in range(stages[s]["num_m"]))

    def ready(q):
        s, m = tiles[q]
        st = stages[s]
        if st["dep"] is not None:
            if st["all_blocks"]:
                if not stage_complete(st["dep"]):
                    return False
            elif not block_complete(st["dep"], m):
                return False
        if st["dep2"] is not None and not stage_complete(st["dep2"]):
            return False
        return True

    n = min(workers, total)
    cursor = list(range(n))                           # the tile each worker is on
    finished = 0
    progress = True
    while progress:
        progress = False
        for w in range(n):
            q = cursor[w]
            if q < total and ready(q):
                s, m = tiles[q]
                done[s][m] += 1
                cursor[w] = q + n
                finished += 1
                progress = True
    return finished, total


SHAPES = [(128, 784, 500), (256, 784, 500), (256, 500, 2000), (4096, 4096, 4096), (1024, 16384, 8192), (96, 130, 72),
          (300, 333, 1000)]


@pytest.mark.parametrize("batch,V,H", SHAPES)
def test_every_tile_of_the_chain_completes(batch, V, H):
    for k, pcd, with_dw, small in itertools.product((1, 2, 3, 10), (False, True), (False, True), (False, True)):
        if small and batch > 512:
            continue                                  # the small-tile variant is chosen for minibatches <= 512 rows
        stages = chain_plan(batch, V, H, k, pcd, with_dw, small)
        for workers in (1, 2, 3, 7, 74, 148):
            finished, total = run(stages, workers)
            assert finished == total, (batch, V, H, k, pcd, with_dw, small, workers, finished, total)


def test_dependencies_point_backwards():
    """The invariant the header states: a tile only ever waits for tiles of earlier stages."""
    for batch, V, H in SHAPES:
        for k, pcd, with_dw in itertools.product((1, 4, 31), (False, True), (False, True)):
            stages = chain_plan(batch, V, H, k, pcd, with_dw)
            assert len(stages) <= 66                  # kMaxChainStages
            for s, st in enumerate(stages):
                assert st["dep"] is None or st["dep"] < s
                assert st["dep2"] is None or st["dep2"] < s


def test_a_forward_dependency_would_deadlock():
    """The model does detect an unsound table: a stage that waits for a later one never completes."""
    stages = chain_plan(256, 784, 500, 2)
    stages[1]["dep"] = 3
    finished, total = run(stages, 4)
    assert finished < total


def test_c3_tile_count():
    """21 projections of 16 x 16 tiles of 256 x 256 = 5376 tiles = 72.6 waves of 74 CTA pairs."""
    stages = chain_plan(4096, 4096, 4096, 10)
    assert len(stages) == 21 and sum(st["num_m"] * st["num_n"] for st in stages) == 5376
