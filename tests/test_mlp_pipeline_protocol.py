"""Model check (CPU) of the mbarrier protocol of fepe_mlp_gemm_persist_kernel (csrc/fepe_mlp.cu): the TMA producer,
the MMA issuer, the operand-transform warps (fused-norm variants) and the epilogue warps are replayed as coroutines
over a faithful model of mbarrier phases (arrival counts, parity waits), with the scheduler picking a random runnable
role at every step.  Checked for every variant and for tile / k-block / stage counts beyond what the GPU tests run:
no deadlock, every stage is transformed after it has landed and before it is multiplied, no stage is overwritten
before the MMA has read it, no TMEM accumulator buffer is overwritten before the epilogue has drained it, and every
tile is stored exactly once.  The parities below are copied from the kernel; if the kernel's protocol changes, this
file has to change with it."""
import random

import pytest


class MBar:
    """mbarrier with `count` expected arrivals per phase (transaction bytes are folded into the arrival)."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        """try_wait.parity: true once the phase with this parity has completed."""
        return self.phase != parity


def run(n_local, num_kb, stages, fuse, seed):
    ng = 1 if fuse == 2 else 2                     # epilogue groups of four warps
    tw = 0 if fuse == 0 else (8 if fuse == 2 else 4)
    full = [MBar(1) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    ready = [MBar(max(tw, 1)) for _ in range(stages)]
    tmem_full = [MBar(1) for _ in range(2)]
    tmem_empty = [MBar(4 * ng) for _ in range(2)]
    stage_state = ["free"] * stages                # free -> landed -> transformed(tw times) -> read -> free
    stage_owner = [None] * stages
    transformed = [0] * stages
    acc_state = ["free", "free"]                   # free -> accumulating -> complete -> (drained by 4 ng warps) -> free
    acc_tile = [None, None]
    drained = [0, 0]
    stored = {}

    def producer():
        it = 0
        for j in range(n_local):
            for kb in range(num_kb):
                s, ph = it % stages, (it // stages) & 1
                while not empty[s].passed(ph ^ 1):
                    yield
                assert stage_state[s] == "free", "TMA overwrites a stage the MMA has not read"
                stage_state[s], stage_owner[s], transformed[s] = "landed", (j, kb), 0
                full[s].arrive()                    # expect_tx + complete_tx of the copies
                it += 1
                yield

    def transform(w):
        it = 0
        for j in range(n_local):
            for kb in range(num_kb):
                s, ph = it % stages, (it // stages) & 1
                while not full[s].passed(ph):
                    yield
                assert stage_state[s] == "landed" and stage_owner[s] == (j, kb)
                transformed[s] += 1
                ready[s].arrive()
                it += 1
                yield

    def mma():
        it = 0
        for j in range(n_local):
            b, use = j & 1, j >> 1
            while not tmem_empty[b].passed((use & 1) ^ 1):
                yield
            assert acc_state[b] == "free", "MMA overwrites an accumulator the epilogue has not drained"
            acc_state[b], acc_tile[b] = "accumulating", j
            for kb in range(num_kb):
                s, ph = it % stages, (it // stages) & 1
                bar = ready[s] if fuse else full[s]
                while not bar.passed(ph):
                    yield
                assert stage_state[s] == "landed" and stage_owner[s] == (j, kb)
                assert transformed[s] == tw, "MMA reads a stage before every transform warp has finished it"
                stage_state[s] = "free"              # tcgen05.commit -> empty[s] once the MMAs have read the stage
                empty[s].arrive()
                if kb == num_kb - 1:
                    acc_state[b] = "complete"
                    tmem_full[b].arrive()
                it += 1
                yield

    def epilogue(w):
        for j in range(n_local):
            b, use = j & 1, j >> 1
            while not tmem_full[b].passed(use & 1):
                yield
            assert acc_state[b] == "complete" and acc_tile[b] == j
            yield                                    # TMEM -> registers -> tile, statistics, stores of all passes
            stored[(j, w)] = stored.get((j, w), 0) + 1
            drained[b] += 1
            if drained[b] == 4 * ng:
                drained[b], acc_state[b] = 0, "free"
            tmem_empty[b].arrive()
            yield

    roles = [producer(), mma()] + [transform(w) for w in range(tw)] + [epilogue(w) for w in range(4 * ng)]
    rng = random.Random(seed)
    live = list(range(len(roles)))
    idle_rounds = 0
    while live:
        progressed = False
        for i in rng.sample(live, len(live)):
            before = (tuple(b.phase for b in full + empty + ready + tmem_full + tmem_empty),
                      tuple(b.pending for b in full + empty + ready + tmem_full + tmem_empty), len(stored))
            try:
                next(roles[i])
            except StopIteration:
                live.remove(i)
                progressed = True
                continue
            after = (tuple(b.phase for b in full + empty + ready + tmem_full + tmem_empty),
                     tuple(b.pending for b in full + empty + ready + tmem_full + tmem_empty), len(stored))
            progressed |= before != after
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 4, f"deadlock: roles {live} wait forever"
    assert all(stored.get((j, w), 0) == 1 for j in range(n_local) for w in range(4 * ng)), "a tile was not stored once"


@pytest.mark.parametrize("fuse", [0, 1, 2])
@pytest.mark.parametrize("stages", [2, 4, 6])
@pytest.mark.parametrize("num_kb", [1, 2, 3, 16])
@pytest.mark.parametrize("n_local", [1, 2, 3, 7])
def test_persistent_gemm_barrier_protocol(n_local, num_kb, stages, fuse):
    for seed in range(8):
        run(n_local, num_kb, stages, fuse, seed)


def test_model_detects_a_wrong_parity():
    """The checker is not vacuous: an MMA that waits for the wrong parity of tmem_empty deadlocks or trips an assertion."""
    import types
    src = open(__file__).read().replace("passed((use & 1) ^ 1)", "passed(use & 1)")
    bad = types.ModuleType("bad_protocol")
    exec(compile(src, "bad_protocol", "exec"), bad.__dict__)
    with pytest.raises(AssertionError):
        for seed in range(8):
            bad.run(3, 2, 4, 0, seed)
