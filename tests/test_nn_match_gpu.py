"""GPU parity of fepe_nn_match (mutual nearest-neighbour descriptor matching) against the numpy
restatement of PointTracker.nn_match_two_way (oracle/nn_match_oracle.py; the `superpoint` package is not part of the
reference tree, so parity is pinned to its published algorithm only).  Bit-exact on descriptors whose dot products
are exact in fp32 (that also exercises argmin's first-occurrence rule on ties); on generic unit descriptors the
decisions may differ only where the two best distances are within rounding of each other."""
import numpy as np
import pytest
import torch

from fepe_b200 import ops
from fepe_b200.matching import crop_or_pad_choice_batch, get_matches_from_descriptors
from oracle import nn_match_oracle as NO

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _grid_descriptors(rng, n, D=256, nnz=16):
    """unit descriptors with `nnz` entries of +-1/4: every dot product is a multiple of 1/16, exact in any order"""
    d = np.zeros((n, D), dtype=np.float32)
    for i in range(n):
        d[i, rng.choice(D, nnz, replace=False)] = rng.choice([-0.25, 0.25], nnz)
    return d


def _check_exact(d1, d2, n1, n2, thresh):
    B = d1.shape[0]
    i1, i2, sc, cnt = ops.nn_match_two_way(T(d1).cuda(), T(d2).cuda(), thresh,
                                           T(n1).cuda() if n1 is not None else None, T(n2).cuda() if n2 is not None else None)
    torch.cuda.synchronize()
    for b in range(B):
        a, c = (d1.shape[1] if n1 is None else n1[b]), (d2.shape[1] if n2 is None else n2[b])
        ref = NO.nn_match_two_way(d1[b, :a].T, d2[b, :c].T, thresh)
        n = int(cnt[b])
        assert n == ref.shape[1], (b, n, ref.shape[1])
        np.testing.assert_array_equal(i1[b, :n].cpu().numpy(), ref[0].astype(np.int32))
        np.testing.assert_array_equal(i2[b, :n].cpu().numpy(), ref[1].astype(np.int32))
        np.testing.assert_array_equal(sc[b, :n].cpu().numpy(), ref[2].astype(np.float32))


@pytest.fixture(params=["simt", "tc"])
def dist_kernel(request, dispatch):
    """every matcher test runs on both distance kernels: plain fp32 on CUDA cores and split-fp16 on tcgen05 (the default)"""
    dispatch("nn_dist", request.param)
    return request.param


@pytest.mark.parametrize("B,N1,N2,thresh", [(3, 300, 257, 0.7), (2, 1000, 1200, 1.0), (4, 64, 500, 1.2), (1, 129, 128, 2.5)])
def test_bit_exact_on_exactly_representable_descriptors(B, N1, N2, thresh, dist_kernel):
    rng = np.random.default_rng(N1 + N2)
    d1 = np.stack([_grid_descriptors(rng, N1) for _ in range(B)])
    d2 = np.stack([_grid_descriptors(rng, N2) for _ in range(B)])
    for b in range(B):                      # plant true correspondences (and exact duplicates -> ties)
        k = min(N1, N2) // 2
        d2[b, rng.permutation(N2)[:k]] = d1[b, rng.permutation(N1)[:k]]
        d2[b, 5] = d2[b, 3]
    _check_exact(d1, d2, None, None, thresh)
    n1 = rng.integers(1, N1 + 1, size=B).astype(np.int32)
    n2 = rng.integers(1, N2 + 1, size=B).astype(np.int32)
    _check_exact(d1, d2, n1, n2, thresh)


def test_generic_unit_descriptors_agree_up_to_rounding_ties(dist_kernel):
    # scores: fp32 CUDA cores differ from numpy by summation order only (2e-6 on the distance); the tensor-core kernel
    # adds the split's 3e-8 and the tensor memory's truncating accumulation (~1.5e-6 on a dot product near 1, which
    # sqrt(2 - 2 dot) magnifies to ~7e-6 on a distance of 0.3: CPU emulation of the arithmetic)
    score_tol = 2e-6 if dist_kernel == "simt" else 2e-5
    rng = np.random.default_rng(5)
    B, N1, N2, D = 2, 800, 900, 256
    d1 = rng.normal(size=(B, N1, D)).astype(np.float32)
    d2 = rng.normal(size=(B, N2, D)).astype(np.float32)
    d2[:, :400] = d1[:, 100:500] + 0.3 * rng.normal(size=(B, 400, D)).astype(np.float32)     # noisy true matches
    d1 /= np.linalg.norm(d1, axis=2, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=2, keepdims=True)
    i1, i2, sc, cnt = ops.nn_match_two_way(T(d1).cuda(), T(d2).cuda(), 1.0)
    for b in range(B):
        ref = NO.nn_match_two_way(d1[b].T, d2[b].T, 1.0)
        n = int(cnt[b])
        ours = set(zip(i1[b, :n].tolist(), i2[b, :n].tolist()))
        theirs = set(zip(ref[0].astype(int).tolist(), ref[1].astype(int).tolist()))
        assert len(ours ^ theirs) <= max(2, len(theirs) // 200), (len(ours), len(theirs), len(ours ^ theirs))
        assert len(theirs) >= 350                                          # the planted matches are found
        both = {p: s for p, s in zip(zip(ref[0].astype(int).tolist(), ref[1].astype(int).tolist()), ref[2])}
        for k in range(n):
            p = (int(i1[b, k]), int(i2[b, k]))
            if p in both:
                assert abs(float(sc[b, k]) - both[p]) < score_tol
        assert (np.diff(i1[b, :n].cpu().numpy()) > 0).all()                # ordered by the first index


def test_match_construction_invariants():
    """matching.get_matches_from_descriptors: the reference's dict (train_good_utils.py:717-724) with the random
    crop / pad of utils_misc.crop_or_pad_choice -- not reproducible by construction, so its invariants are checked."""
    rng = np.random.default_rng(9)
    B, N, D, out = 3, 400, 256, 150
    d1 = np.stack([_grid_descriptors(rng, N) for _ in range(B)])
    d2 = d1[:, rng.permutation(N)].copy()
    d2[1, 100:] = np.stack([_grid_descriptors(rng, 1) for _ in range(300)])[:, 0]     # sample 1: only ~100 true matches
    pts1 = rng.uniform(0, 1000, size=(B, N, 2)).astype(np.float32)
    pts2 = rng.uniform(0, 1000, size=(B, N, 2)).astype(np.float32)
    g = torch.Generator(device="cuda").manual_seed(1)
    r = get_matches_from_descriptors(T(pts1).cuda(), T(pts2).cuda(), T(d1).cuda(), T(d2).cuda(), 0.7, out, generator=g)
    assert r["xs"].shape == (B, out, 4) and r["quality"].shape == (B, out, 1) and r["offsets"].shape == (B, out, 4)
    for b in range(B):
        ref = NO.nn_match_two_way(d1[b].T, d2[b].T, 0.7)
        n = ref.shape[1]
        assert int(r["num_matches"][b]) == n
        valid = {(tuple(pts1[b, int(i)]), tuple(pts2[b, int(j)])): s for i, j, s in ref.T}
        xs = r["xs"][b].cpu().numpy()
        rows = [(tuple(x[:2]), tuple(x[2:])) for x in xs]
        assert all(p in valid for p in rows)                                # every row is a true mutual match
        q = r["quality"][b, :, 0].cpu().numpy()
        assert all(abs(valid[p] - qq) < 1e-6 for p, qq in zip(rows, q))
        if n >= out:
            assert len(set(rows)) == out                                    # a permutation prefix: no duplicates
        else:
            assert len(set(rows[:n])) == n and set(rows) == set(valid)      # all matches once, then a resample
    c = crop_or_pad_choice_batch(torch.tensor([0, 5, 400], device="cuda"), 400, 10, g)
    assert c.shape == (3, 10) and int(c[1].max()) < 5 and int(c[0].max()) == 0
