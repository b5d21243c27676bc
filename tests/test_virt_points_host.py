"""CPU checks of the ground-truth / virtual-point construction:
  * the numpy oracle (oracle/virt_points_oracle.py) against cv2.correctMatches itself and against the committed
    outputs of the reference's own E_F_from_Rt_np / get_virt_x1x2_np / R_to_q_np (tests/golden/gt_virt_ref.npz);
  * the product's device math (csrc/fepe_virt.cuh compiled for the host) against both."""
import ctypes
import os

import numpy as np
import pytest

from oracle import virt_points_oracle as VO

try:
    import cv2
except Exception:       # pragma: no cover - cv2 is part of the image
    cv2 = None

PX_TOL = 2e-3           # pixels; float32 points at ~1e3 px carry 6e-5 px of rounding, F built from float32 K / Rt 1e-7 relative
DP, FP = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def ref():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "gt_virt_ref.npz"), allow_pickle=False))


def _shim_correct(shim, F, p1, p2):
    F64 = np.ascontiguousarray(np.asarray(F, dtype=np.float64).reshape(-1))
    p1 = np.ascontiguousarray(p1, dtype=np.float32)
    p2 = np.ascontiguousarray(p2, dtype=np.float32)
    o1, o2 = np.zeros_like(p1), np.zeros_like(p2)
    nan = shim.shim_correct_matches(F64.ctypes.data_as(DP), p1.ctypes.data_as(FP), p2.ctypes.data_as(FP), p1.shape[0],
                                    o1.ctypes.data_as(FP), o2.ctypes.data_as(FP))
    return o1, o2, nan


def _scrub(a):
    a = np.array(a, copy=True)
    a[np.isnan(a)] = 0.0
    return a


def test_oracle_matches_committed_cv2_outputs(ref):
    """Raw cv2.correctMatches outputs for random rank-2 F at three scalings (full-degree and truncated polynomials,
    NaN cases included)."""
    for F, p1, p2, c1, c2 in zip(ref["cv_F"], ref["cv_p1"], ref["cv_p2"], ref["cv_c1"], ref["cv_c2"]):
        o1, o2 = VO.correct_matches(F, p1, p2)
        assert (np.isnan(o1) == np.isnan(c1)).all() and (np.isnan(o2) == np.isnan(c2)).all()
        scale = max(1.0, float(np.nanmax(np.abs(c1))))
        assert np.nanmax(np.abs(o1 - c1)) <= 1e-5 * scale and np.nanmax(np.abs(o2 - c2)) <= 1e-5 * scale


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_oracle_matches_cv2_live():
    rng = np.random.default_rng(5)
    from fepe_b200 import synth
    g1, g2 = VO.virt_grid(synth.KITTI_IMAGE_SIZE)
    d = synth.make_batch(6, 8, seed=21)
    for b in range(6):                                        # pixel-unit F of KITTI-like scenes: truncated polynomials
        F = d["F_gt"][b].astype(np.float64)
        c1, c2 = cv2.correctMatches(F, g2[None], g1[None])
        o1, o2 = VO.correct_matches(F, g2, g1)
        assert (np.isnan(o1) == np.isnan(c1[0])).all()
        assert np.nanmax(np.abs(o1 - c1[0])) <= 1e-4 and np.nanmax(np.abs(o2 - c2[0])) <= 1e-4
    for _ in range(6):                                        # normalised F: all six roots
        U, S, Vt = np.linalg.svd(rng.normal(size=(3, 3)))
        F = U @ np.diag([S[0], S[1], 0.0]) @ Vt
        p1, p2 = rng.uniform(-1, 1, size=(30, 2)), rng.uniform(-1, 1, size=(30, 2))
        c1, c2 = cv2.correctMatches(F, p1[None], p2[None])
        o1, o2 = VO.correct_matches(F, p1, p2)
        assert (np.isnan(o1) == np.isnan(c1[0])).all()
        assert np.nanmax(np.abs(o1 - c1[0])) <= 1e-9 and np.nanmax(np.abs(o2 - c2[0])) <= 1e-9


def test_oracle_matches_reference_sample_keys(ref):
    """oracle.gt_sample == the reference's dataset code on the same float32 inputs, key by key."""
    grids = (ref["grid1"], ref["grid2"])
    zeroed = 0
    for i in range(ref["K"].shape[0]):
        g = VO.gt_sample(ref["Rt"][i], ref["K"][i], None, grids)
        for k in ("E", "F", "q_cam", "t_cam", "q_scene", "t_scene"):
            assert g[k].shape == ref[k][i].shape
            np.testing.assert_allclose(g[k], ref[k][i], rtol=1e-6, atol=1e-9, err_msg=k)
        for k in ("pts1_virt", "pts2_virt"):
            np.testing.assert_allclose(g[k], ref[k][i], rtol=0, atol=1e-4, err_msg=k)
        for k in ("pts1_virt_normalized", "pts2_virt_normalized"):
            np.testing.assert_allclose(g[k], ref[k][i], rtol=0, atol=1e-6, err_msg=k)
        zeroed += int((ref["pts1_virt"][i][:, :2] == 0).all(-1).sum())
    assert zeroed >= 1                                       # the fixture holds a NaN -> 0 point


def test_virtual_points_satisfy_the_epipolar_constraint(ref):
    """What the loss relies on (utils_misc.py:174 'SHOULD BE ALL ZEROS'): x2^T F x1 = 0 for every corrected pair."""
    for i in range(ref["K"].shape[0]):
        g = VO.gt_sample(ref["Rt"][i], ref["K"][i], None, (ref["grid1"], ref["grid2"]))
        keep = ~(g["pts1_virt"][:, :2] == 0).all(-1)
        F = g["F"].astype(np.float64)
        l2 = g["pts1_virt"][keep].astype(np.float64) @ F.T
        d = np.abs(np.sum(l2 * g["pts2_virt"][keep], axis=1)) / np.linalg.norm(l2[:, :2], axis=1)
        assert d.max() < 5e-3                                 # pixels; float32 coordinates


def test_device_math_solve_poly_matches_oracle(shim):
    rng = np.random.default_rng(2)
    for trial in range(40):
        k = rng.normal(size=7) * 10.0 ** rng.integers(-3, 2, size=7)
        if trial % 4 == 1:
            k[6] = 1e-17                                      # degree drops to 5
        if trial % 4 == 2:
            k[3:] = rng.normal(size=4) * 1e-18                # degree drops to 2
        if trial % 4 == 3:
            k[2:] = rng.normal(size=5) * 1e-18                # degree drops to 1
        want, n_want = VO.solve_poly(k)
        got = np.zeros(6)
        n = shim.shim_solve_poly6(np.ascontiguousarray(k).ctypes.data_as(DP), got.ctypes.data_as(DP))
        assert n == n_want
        np.testing.assert_allclose(got, np.array(want), rtol=1e-9, atol=1e-12)


def test_device_math_correct_matches(shim, ref):
    """fepe_virt.cuh (host build) against the committed cv2 outputs and the reference's virtual points."""
    for F, p1, p2, c1, c2 in zip(ref["cv_F"], ref["cv_p1"], ref["cv_p2"], ref["cv_c1"], ref["cv_c2"]):
        o1, o2, nan = _shim_correct(shim, F, p1, p2)
        assert nan == int(np.isnan(c1[:, 0]).sum())
        scale = max(1.0, float(np.nanmax(np.abs(c1))))
        assert np.abs(o1 - _scrub(c1)).max() <= 1e-5 * scale and np.abs(o2 - _scrub(c2)).max() <= 1e-5 * scale
    for i in range(ref["K"].shape[0]):
        o1, o2, _ = _shim_correct(shim, ref["F"][i], ref["grid2"], ref["grid1"])     # the reference's argument order
        assert np.abs(o1 - ref["pts1_virt"][i][:, :2]).max() <= PX_TOL
        assert np.abs(o2 - ref["pts2_virt"][i][:, :2]).max() <= PX_TOL


def test_device_math_gt_from_motion(shim, ref):
    for i in range(ref["K"].shape[0]):
        K = np.ascontiguousarray(ref["K"][i], dtype=np.float32)
        Rt = np.ascontiguousarray(ref["Rt"][i], dtype=np.float32)
        gt = np.zeros(32)
        shim.shim_gt_from_motion(K.ctypes.data_as(FP), Rt.ctypes.data_as(FP), gt.ctypes.data_as(DP))
        np.testing.assert_allclose(gt[0:9].reshape(3, 3), ref["E"][i], rtol=0, atol=2e-6)
        # the reference forms K^-T E K^-1 in float32: O(1) intermediate terms cancel down to |F| ~ 1e-2, leaving
        # ~1e-7 of absolute float32 rounding in its own F; the kernel works in fp64
        np.testing.assert_allclose(gt[9:18].reshape(3, 3), ref["F"][i], rtol=0, atol=5e-7)
        np.testing.assert_allclose(gt[18:22], ref["q_cam"][i][:, 0], rtol=0, atol=2e-6)
        np.testing.assert_allclose(gt[22:25], ref["t_cam"][i][:, 0], rtol=0, atol=2e-6)
        np.testing.assert_allclose(gt[25:29], ref["q_scene"][i][:, 0], rtol=0, atol=2e-6)
        np.testing.assert_allclose(gt[29:32], ref["t_scene"][i][:, 0], rtol=0, atol=2e-6)
