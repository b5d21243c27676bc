"""Model check (CPU) of the self-refilling stage ring of fepe_gram_kernel (csrc/fepe_fit_split.cu, K1 of the split
pipeline): S shared-memory stages, G teams, NO producer warp and NO `empty` barrier -- the team that finishes the pair in
a stage issues the bulk copy of the pair that uses the stage next, and the team that owns that pair waits on the stage's
`full` mbarrier with parity (use & 1).

A parity wait tells only two consecutive phases apart, so parity ALONE is correct only while no team reaches the wait
for use k of a stage before the copy of use k-1 has landed.  This discrete-event model (random per-pair durations, random
copy latencies) shows where that breaks: with S = 11, G = 8 never while pair durations differ by less than 1.5x, from
1.75x on in a few per cent of the runs, and always under a long stall of one team.  The kernel therefore carries a
per-stage USE COUNTER: the refilling team writes "use k is on its way" before it issues the copy, and the waiter spins on
that counter before it trusts the parity.  Checked here for both variants: no early pass with the counter under any
schedule (4x spread, long stalls), no deadlock, every pair processed exactly once out of the stage that holds ITS data;
and the parity-only variant is kept as the negative control that reproduces the hazard.
The parities, the refill rule and the counter are copied from the kernel; if the kernel's protocol changes, this file
changes too."""
import heapq
import random

import pytest


def _run(n_local, S, G, seed, stall=None, use_counter=False, spread=4.0):
    """Event-driven replay of one CTA.  Returns (early_passes, processed): early_passes counts waits that the parity rule
    let through although the stage did not hold the waiter's pair; processed = [(pair, pair whose data the stage held)].
    `stall` = (team, local pair, extra time in pair-times)."""
    rng = random.Random(seed)
    completed, holds, issued = [0] * S, [None] * S, [0] * S
    events, seq = [], [0]

    def push(t, kind, payload):
        seq[0] += 1
        heapq.heappush(events, (t, seq[0], kind, payload))

    def fill(t, j):
        s = j % S
        issued[s] += 1
        holds[s] = None
        push(t + rng.uniform(0.05, 0.6), "land", j)

    for j in range(min(S, n_local)):
        fill(0.0, j)
    next_pair = list(range(G))
    early, processed = 0, []
    waiting = [[] for _ in range(S)]    # teams spinning on the stage's barrier: re-tried whenever its state changes
    for g in range(G):
        push(0.0, "try", g)

    def wake(t, s):
        for g in waiting[s]:
            push(t, "try", g)
        waiting[s].clear()

    while events:
        t, _, kind, x = heapq.heappop(events)
        if kind == "land":
            completed[x % S] += 1
            holds[x % S] = x
            wake(t, x % S)
        elif kind == "refill":
            fill(t, x)
            wake(t, x % S)
        else:
            g = x
            j = next_pair[g]
            if j >= n_local:
                continue
            s, k = j % S, j // S
            passes = (completed[s] & 1) != (k & 1)            # try_wait.parity(k & 1)
            if use_counter:
                passes = passes and issued[s] == k + 1        # the refilling team has announced use k of this stage
            if not passes:
                waiting[s].append(g)
                continue
            if holds[s] != j:
                early += 1
            processed.append((j, holds[s]))
            dur = rng.uniform(1.0, spread)
            if stall is not None and stall[0] == g and stall[1] == j:
                dur += stall[2]
            if j + S < n_local:
                push(t + dur, "refill", j + S)
            next_pair[g] = j + G
            push(t + dur, "try", g)
    assert all(not w for w in waiting), "deadlock: a team waits for a phase that never completes"
    return early, processed


@pytest.mark.parametrize("S,G", [(11, 8), (10, 6), (5, 4), (13, 8), (11, 4)])
@pytest.mark.parametrize("n_local", [1, 7, 64, 221, 2000])
def test_use_counter_never_passes_a_wait_early(S, G, n_local):
    """The shipped protocol (parity + use counter) under durations spread over a factor of four."""
    for seed in range(12 if n_local > 1000 else 40):
        early, processed = _run(n_local, S, G, seed, use_counter=True)
        assert early == 0, (S, G, n_local, seed)
        assert sorted(j for j, _ in processed) == list(range(n_local))          # every pair exactly once
        assert all(j == h for j, h in processed)                                  # out of the stage holding ITS data


@pytest.mark.parametrize("S,G", [(11, 8), (10, 6), (13, 8)])
def test_parity_alone_is_safe_only_for_similar_pair_durations(S, G):
    """Negative control: without the counter the ring is fine for durations within 1.5x of each other and breaks beyond."""
    for seed in range(60):
        early, processed = _run(221, S, G, seed, spread=1.5)
        assert early == 0, (S, G, seed)
        assert all(j == h for j, h in processed)
    broken = 0
    for seed in range(200):
        try:
            early, _ = _run(221, S, G, seed, spread=4.0)
        except AssertionError:          # a wrong pass can also leave a team waiting for a phase that never comes
            early = 1
        broken += early > 0
    assert broken > 0, "the model no longer reproduces the hazard of a parity-only ring"


def test_a_long_stall_breaks_parity_alone_and_not_the_use_counter():
    S, G, n_local = 11, 8, 96
    broken = 0
    for seed in range(30):
        try:
            early, _ = _run(n_local, S, G, seed, stall=(3, 19, 40.0))
        except AssertionError:
            early = 1
        broken += early > 0
        early_fixed, processed = _run(n_local, S, G, seed, stall=(3, 19, 40.0), use_counter=True)
        assert early_fixed == 0
        assert sorted(j for j, _ in processed) == list(range(n_local))
    assert broken > 0, "the model no longer reproduces the hazard the use counter exists for"
