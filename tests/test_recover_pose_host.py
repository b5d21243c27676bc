"""CPU checks of the validation pose recovery:
  * the numpy oracle (oracle/recover_pose_oracle.py) against cv2.recoverPose itself and against the committed outputs
    of the reference's own goodCorr_eval_nondecompose (tests/golden/recover_pose_ref.npz);
  * the product's device math (csrc/fepe_recover.cuh compiled for the host) against both."""
import ctypes

import numpy as np
import pytest

from fepe_b200 import synth
from oracle import recover_pose_oracle as RO

try:
    import cv2
except Exception:       # pragma: no cover - cv2 is part of the image
    cv2 = None


@pytest.fixture(scope="module")
def ref():
    import os
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "recover_pose_ref.npz"), allow_pickle=False))


def _run_shim(shim, E, K, m, Rt, thresh=50.0):
    dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    E64 = np.ascontiguousarray(E.astype(np.float32).astype(np.float64).reshape(-1))     # the kernel reads fp32 E
    K64 = np.ascontiguousarray(K.astype(np.float32).astype(np.float64).reshape(-1))
    m32 = np.ascontiguousarray(m.astype(np.float32))
    Rt64 = np.ascontiguousarray(Rt.astype(np.float32).astype(np.float64).reshape(-1))
    N = m32.shape[0]
    R, t, errs = np.zeros(9), np.zeros(3), np.zeros(2)
    counts, best = np.zeros(4, dtype=np.int32), np.zeros(1, dtype=np.int32)
    mask = np.zeros(N, dtype=np.uint8)
    shim.shim_recover_pose(E64.ctypes.data_as(dp), K64.ctypes.data_as(dp), m32.ctypes.data_as(fp), N, thresh,
                           Rt64.ctypes.data_as(dp), R.ctypes.data_as(dp), t.ctypes.data_as(dp),
                           counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), best.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                           errs.ctypes.data_as(dp), mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    return R.reshape(3, 3), t, counts, int(best[0]), errs, mask


def test_oracle_matches_reference_function_outputs(ref):
    """goodCorr_eval_nondecompose(x1, x2, E, inv(Rt)[:3], K, None) of the unmodified reference -> (M, (err_q, err_t))."""
    for i in range(ref["E"].shape[0]):
        K, m = ref["K"][i].astype(np.float64), ref["matches"][i]
        good, R, t, mask, _ = RO.recover_pose(ref["E"][i].astype(np.float64), m[:, :2], m[:, 2:], K[0, 0], (K[0, 2], K[1, 2]))
        assert good == int(ref["good"][i])
        np.testing.assert_allclose(np.hstack((R, t[:, None])), ref["M"][i], atol=1e-9)
        np.testing.assert_array_equal(mask, ref["mask"][i] > 0)
        eq, et = RO.pose_errors_vs_gt(R, t, ref["Rt"][i])
        assert abs(eq - ref["err"][i][0]) < 1e-5 and abs(et - ref["err"][i][1]) < 1e-5      # degrees (near 0, d acos is steep)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_oracle_matches_cv2_recover_pose_live():
    rng = np.random.default_rng(3)
    for seed in range(5):
        d = synth.make_batch(3, 257, seed=40 + seed)
        for b in range(3):
            K, m = d["Ks"][b].astype(np.float64), d["matches_xy_ori"][b]
            E = d["E_gt"][b].astype(np.float64) + (0.02 if b else 0.0) * rng.normal(size=(3, 3))
            n, R, t, mask = cv2.recoverPose(E, m[:, :2].astype(np.float64), m[:, 2:].astype(np.float64),
                                            focal=float(K[0, 0]), pp=(float(K[0, 2]), float(K[1, 2])))
            good, Ro, to, mo, _ = RO.recover_pose(E, m[:, :2], m[:, 2:], K[0, 0], (K[0, 2], K[1, 2]))
            assert good == n
            np.testing.assert_allclose(Ro, R, atol=1e-12)
            np.testing.assert_allclose(to, t.reshape(-1), atol=1e-12)
            np.testing.assert_array_equal(mo, mask.reshape(-1) > 0)


def test_device_math_matches_reference_outputs(shim, ref):
    """fepe_recover.cuh on the host (Jacobi 4x4 DLT, cheirality, selection, error metrics) vs the reference's outputs.
    The winner, its (R, t), the count and the mask must be identical; angles to 1e-4 degrees."""
    for i in range(ref["E"].shape[0]):
        R, t, counts, best, errs, mask = _run_shim(shim, ref["E"][i], ref["K"][i], ref["matches"][i], ref["Rt"][i])
        np.testing.assert_allclose(np.hstack((R, t[:, None])), ref["M"][i], atol=1e-6)     # E is fp32 on our side
        assert int(counts[best]) == int(ref["good"][i])
        np.testing.assert_array_equal(mask > 0, ref["mask"][i] > 0)
        assert abs(errs[0] - ref["err"][i][0]) < 1e-4 and abs(errs[1] - ref["err"][i][1]) < 1e-4


def test_device_math_counts_match_oracle_per_candidate(shim):
    """All four candidate counts (as a multiset: the candidate ORDER depends on SVD signs) on noisy scenes with far
    points, where the 50-unit distance test and the cheirality test both bite."""
    rng = np.random.default_rng(11)
    for seed in range(4):
        d = synth.make_batch(2, 500, seed=70 + seed)
        for b in range(2):
            K, m = d["Ks"][b], d["matches_xy_ori"][b]
            E = (d["E_gt"][b] + 0.05 * rng.normal(size=(3, 3))).astype(np.float32)
            R, t, counts, best, errs, mask = _run_shim(shim, E, K, m, d["delta_Rtijs_4_4"][b])
            K64 = K.astype(np.float64)
            good, Ro, to, mo, cnt = RO.recover_pose(E.astype(np.float64), m[:, :2], m[:, 2:], K64[0, 0], (K64[0, 2], K64[1, 2]))
            assert sorted(counts.tolist()) == sorted(cnt.tolist())
            assert int(counts[best]) == good
            np.testing.assert_allclose(R, Ro, atol=1e-6)
            np.testing.assert_allclose(t, to, atol=1e-6)
            np.testing.assert_array_equal(mask > 0, mo)
