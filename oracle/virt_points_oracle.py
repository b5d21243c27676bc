"""CPU ORACLE for the ground-truth side of a sample -- TEST INFRASTRUCTURE ONLY.

What the reference's dataset does per sample on the host before the hot path's loss can be evaluated
(deepFEPE/datasets/kitti_odo_corr.py):

    E, F = utils_F.E_F_from_Rt_np(Rt_scene[:3,:3], Rt_scene[:3,3:4], K)              # :290-302 -> utils_F.py:835-846
    pts1_virt_normalized, pts2_virt_normalized, pts1_virt, pts2_virt =
        utils_misc.get_virt_x1x2_np(image_size, F, K, pts1_virt_b, pts2_virt_b)       # :526-541 -> utils_misc.py:173-199
    Rt_cam = inv(Rt_scene); q_cam = R_to_q_np(Rt_cam[:3,:3]); t_cam = Rt_cam[:3,3:4]  # :547-554 -> utils_geo.py:88-117
    q_scene = R_to_q_np(Rt_scene[:3,:3]); t_scene = Rt_scene[:3,3:4]

get_virt_x1x2_np moves a 10x10 pixel grid (utils_misc.py:163-171) onto the ground-truth epipolar geometry with
``cv2.correctMatches(F_gt, pts2_virt_b[None], pts1_virt_b[None])`` and zeroes the NaNs OpenCV produces (:176-178).
That arithmetic lives in a THIRD-PARTY dependency that is not under /root/reference: OpenCV (opencv-python 3.4.2.16
pinned in requirements.txt; 4.13.0 in this image), modules/calib3d/src/triangulate.cpp icvCorrectMatches -- the
optimal two-view correction of Hartley & Sturm ("Triangulation", CVIU 1997; Hartley & Zisserman alg. 12.1).  Its
published algorithm, restated below in numpy fp64, per correspondence (x1,y1) <-> (x2,y2) with x2^T F x1 = 0 wanted:

  1. F0 = T2^T F T1 with T = [[1,0,x],[0,1,y],[0,0,1]] (both points moved to the origin);
  2. right / left null vectors e1, e2 of F0 (SVD), each scaled so that ex^2 + ey^2 = 1; a correspondence whose
     epipole lies at infinity (ex = ey = 0) is left unchanged;
  3. F1 = R2 F0 R1^T with R = [[ex,ey,0],[-ey,ex,0],[0,0,1]]; f1 = e1z, f2 = e2z, a = F1[1,1], b = F1[1,2],
     c = F1[2,1], d = F1[2,2];
  4. the six roots of g(t) = t((at+b)^2 + f2^2 (ct+d)^2)^2 - (ad-bc)(1+f1^2 t^2)^2 (at+b)(ct+d) through
     cvSolvePoly(g, roots, 100, 20) = cv::solvePoly (modules/core/src/mathfuncs.cpp), restated in solve_poly():
     leading coefficients with |k| <= DBL_EPSILON are dropped (degree n <= 6), Durand-Kerner from the start values
     (1+i)^j with in-place (Gauss-Seidel) updates, exactly 100 sweeps unless a sweep changes nothing; the 6 - n
     root slots the solver does not compute still hold the coefficient doubles it staged there -- magnitudes
     <= DBL_EPSILON, i.e. t = 0 for every purpose (n = 4, 5 would read uninitialised memory; see below).
     THIS MATTERS: with a fundamental matrix in PIXEL units, as the reference passes it, every coefficient above k1
     or k2 is below DBL_EPSILON, so OpenCV solves a linear or quadratic truncation of g and the corrected points
     are NOT the optimal ones (they differ by up to several pixels); they still satisfy the epipolar constraint
     exactly, which is all the F-loss needs.  Parity means reproducing that, not the textbook optimum;
  5. s(t) = t^2/(1+f1^2 t^2) + (ct+d)^2/((at+b)^2 + f2^2 (ct+d)^2) evaluated at the REAL PART of every root and at
     t = inf (1/f1^2 + c^2/(a^2 + f2^2 c^2)); the smallest wins (the value at infinity only if strictly smaller than
     all six);
  6. closest points to the origin on l1 = (t f1, 1, -t), l2 = (-f2(ct+d), at+b, ct+d), mapped back by T R^T.
     When t = inf wins OpenCV evaluates these with t = DBL_MAX and returns NaN for both points; the reference then
     writes 0 (utils_misc.py:177-178), and so does this oracle.

PARITY PIN: tests/test_virt_points_host.py checks correct_matches() against cv2.correctMatches itself (cv2 is part of
this image) and the whole sample construction against the committed fixture tests/golden/gt_virt_ref.npz, generated
by tests/golden/make_golden_virt.py from the UNMODIFIED reference functions named above.
"""
from __future__ import annotations

import numpy as np


def virt_grid(im_shape, step: float = 0.1):
    """deepFEPE/dsac_tools/utils_misc.py:163-171 get_virt_x1x2_grid -> two identical [100,2] float32 pixel grids."""
    xx, yy = np.meshgrid(np.arange(0, 1, step), np.arange(0, 1, step))
    g = np.float32(np.vstack((im_shape[1] * xx.flatten(), im_shape[0] * yy.flatten())).T)
    return g, g.copy()


def poly_coeffs(a, b, c, d, f1, f2):
    """Coefficients k0..k6 of g(t) (item 4), expanded as icvCorrectMatches does."""
    f1s, f2s = f1 * f1, f2 * f2
    f14, f24 = f1s * f1s, f2s * f2s
    k6 = b * c * c * f14 * a - a * a * d * f14 * c
    k5 = f24 * c ** 4 + 2 * a * a * f2s * c * c - a * a * d * d * f14 + b * b * c * c * f14 + a ** 4
    k4 = (4 * a ** 3 * b + 2 * b * c * c * f1s * a + 4 * f24 * c ** 3 * d + 4 * a * b * f2s * c * c
          + 4 * a * a * f2s * c * d - 2 * a * a * d * f1s * c - a * d * d * f14 * b + b * b * c * f14 * d)
    k3 = (6 * a * a * b * b + 6 * f24 * c * c * d * d + 2 * b * b * f2s * c * c + 2 * a * a * f2s * d * d
          - 2 * a * a * d * d * f1s + 2 * b * b * c * c * f1s + 8 * a * b * f2s * c * d)
    k2 = (4 * a * b ** 3 + 4 * b * b * f2s * c * d + 4 * f24 * c * d ** 3 - a * a * d * c + b * c * c * a
          + 4 * a * b * f2s * d * d - 2 * a * d * d * f1s * b + 2 * b * b * c * f1s * d)
    k1 = f24 * d ** 4 + b ** 4 + 2 * b * b * f2s * d * d - a * a * d * d + b * b * c * c
    k0 = -a * d * d * b + b * b * c * d
    return np.array([k0, k1, k2, k3, k4, k5, k6], dtype=np.float64)


DBL_EPSILON = float(np.finfo(np.float64).eps)


def solve_poly(k, max_iters: int = 100):
    """cv::solvePoly for real coefficients k[0..6] (k[i] multiplies t^i).  Returns (real parts of the 6 output
    slots, degree actually solved).  Complex arithmetic spelled out like OpenCV's Complex<double> operators so the
    rounding sequence is the same (no fused multiply-add)."""
    n0 = len(k) - 1
    n = n0
    while n > 1 and abs(k[n]) <= DBL_EPSILON:
        n -= 1
    re = [0.0] * n
    im = [0.0] * n
    pr, pi = 1.0, 0.0
    for i in range(n):
        re[i], im[i] = pr, pi
        pr, pi = pr - pi, pr + pi                       # p * (1 + i)
    for _ in range(max_iters):
        max_diff = 0.0
        for i in range(n):
            p_r, p_i = re[i], im[i]
            nr, ni = float(k[n]), 0.0
            dr, di = float(k[n]), 0.0
            for j in range(n):
                nr, ni = nr * p_r - ni * p_i + float(k[n - j - 1]), nr * p_i + ni * p_r
                if j != i:
                    qr, qi = p_r - re[j], p_i - im[j]
                    if qr != 0.0 or qi != 0.0:          # coincident roots: OpenCV's multiple-root branch, not restated
                        dr, di = dr * qr - di * qi, dr * qi + di * qr
            t = 1.0 / (dr * dr + di * di)
            nr, ni = (nr * dr + ni * di) * t, (-nr * di + ni * dr) * t
            re[i], im[i] = p_r - nr, p_i - ni
            max_diff = max(max_diff, float(np.sqrt(nr * nr + ni * ni)))
        if max_diff <= 0.0:
            break
    return re + [0.0] * (n0 - n), n


def cost(t, a, b, c, d, f1, f2):
    return t * t / (1 + f1 * f1 * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f2 * f2 * (c * t + d) ** 2)


def correct_pair(F: np.ndarray, x1: float, y1: float, x2: float, y2: float):
    """One correspondence; returns (x1', y1', x2', y2'), NaN where OpenCV returns NaN."""
    T1 = np.array([[1.0, 0.0, x1], [0.0, 1.0, y1], [0.0, 0.0, 1.0]])
    T2 = np.array([[1.0, 0.0, x2], [0.0, 1.0, y2], [0.0, 0.0, 1.0]])
    F0 = T2.T @ F @ T1
    U, _, Vt = np.linalg.svd(F0)
    e1, e2 = Vt[2], U[:, 2]
    n1, n2 = np.hypot(e1[0], e1[1]), np.hypot(e2[0], e2[1])
    if n1 == 0.0 or n2 == 0.0:
        return x1, y1, x2, y2
    e1, e2 = e1 / n1, e2 / n2
    R1 = np.array([[e1[0], e1[1], 0.0], [-e1[1], e1[0], 0.0], [0.0, 0.0, 1.0]])
    R2 = np.array([[e2[0], e2[1], 0.0], [-e2[1], e2[0], 0.0], [0.0, 0.0, 1.0]])
    F1 = R2 @ F0 @ R1.T
    f1, f2, a, b, c, d = e1[2], e2[2], F1[1, 1], F1[1, 2], F1[2, 1], F1[2, 2]
    k = poly_coeffs(a, b, c, d, f1, f2)
    roots, _ = solve_poly(k)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s_val = 1.0 / (f1 * f1) + c * c / (a * a + f2 * f2 * c * c)
        t_min = None
        for t in roots:
            s = cost(t, a, b, c, d, f1, f2)
            if s < s_val:
                s_val, t_min = s, t
    if t_min is None:
        return (np.nan,) * 4
    t = t_min
    q1 = np.array([t * t * f1, t, t * t * f1 * f1 + 1.0])
    q2 = np.array([f2 * (c * t + d) ** 2, -(a * t + b) * (c * t + d), f2 * f2 * (c * t + d) ** 2 + (a * t + b) ** 2])
    p1 = T1 @ R1.T @ (q1 / q1[2])
    p2 = T2 @ R2.T @ (q2 / q2[2])
    return p1[0], p1[1], p2[0], p2[1]


def correct_matches(F: np.ndarray, points1: np.ndarray, points2: np.ndarray):
    """cv2.correctMatches(F, points1[None], points2[None]) for [P,2] point sets: new points with
    points2'^T F points1' = 0 at minimal total squared displacement.  Output dtype follows the points (like cv2)."""
    F = np.asarray(F, dtype=np.float64)
    o1 = np.empty(points1.shape, dtype=np.float64)
    o2 = np.empty(points2.shape, dtype=np.float64)
    for i in range(points1.shape[0]):
        o1[i, 0], o1[i, 1], o2[i, 0], o2[i, 1] = correct_pair(F, float(points1[i, 0]), float(points1[i, 1]),
                                                             float(points2[i, 0]), float(points2[i, 1]))
    return o1.astype(points1.dtype), o2.astype(points2.dtype)


def skew(v: np.ndarray) -> np.ndarray:
    """utils_misc.py:38-46 skew_symmetric_np for a [3,1] vector."""
    x, y, z = float(v[0, 0]), float(v[1, 0]), float(v[2, 0])
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]], dtype=v.dtype)


def E_F_from_Rt(R: np.ndarray, t: np.ndarray, K: np.ndarray):
    """utils_F.py:835-846 E_F_from_Rt_np: E = [t]x R, F = K^-T E K^-1."""
    E = skew(t) @ R
    Ki = np.linalg.inv(K)
    return E, Ki.T @ E @ Ki


def R_to_q(matrix: np.ndarray) -> np.ndarray:
    """utils_geo.py:88-117 R_to_q_np: four-branch trace method on m = R^T, float32 result [4,1] with q[0] >= 0."""
    m = matrix.conj().transpose()
    if m[2, 2] < 0:
        if m[0, 0] > m[1, 1]:
            t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
            q = [m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]]
        else:
            t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
            q = [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]]
    else:
        if m[0, 0] < -m[1, 1]:
            t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
            q = [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t]
        else:
            t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
            q = [t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]]
    q = np.array(q, dtype=np.float32)
    q *= 0.5 / np.sqrt(t)
    if q[0] < 0.0:
        q = -q
    return q.reshape(-1, 1)


def virt_x1x2(F_gt: np.ndarray, K: np.ndarray, pts1_virt_b: np.ndarray, pts2_virt_b: np.ndarray):
    """utils_misc.py:173-199 get_virt_x1x2_np.  Note the argument order of the correctMatches call (the second grid
    is OpenCV's points1; both grids are equal) and that BOTH normalised outputs are K^-1 pts1_virt (:197-198)."""
    p1, p2 = correct_matches(F_gt, pts2_virt_b, pts1_virt_b)
    p1[np.isnan(p1)] = 0.0
    p2[np.isnan(p2)] = 0.0
    h1 = np.hstack((p1, np.ones((p1.shape[0], 1), dtype=p1.dtype)))
    h2 = np.hstack((p2, np.ones((p2.shape[0], 1), dtype=p2.dtype)))
    n1 = (np.linalg.inv(K) @ h1.T).T
    return n1, n1.copy(), h1, h2


def gt_sample(Rt_scene: np.ndarray, K: np.ndarray, im_shape, grids=None) -> dict:
    """The ground-truth keys of one sample (kitti_odo_corr.py:290-302, :526-566)."""
    g1, g2 = grids if grids is not None else virt_grid(im_shape)
    E, F = E_F_from_Rt(Rt_scene[:3, :3], Rt_scene[:3, 3:4], K)
    n1, n2, h1, h2 = virt_x1x2(F, K, g1, g2)
    Rt_cam = np.linalg.inv(Rt_scene)
    return {"E": E, "F": F, "pts1_virt_normalized": n1, "pts2_virt_normalized": n2, "pts1_virt": h1, "pts2_virt": h2,
            "q_cam": R_to_q(Rt_cam[:3, :3]), "t_cam": Rt_cam[:3, 3:4],
            "q_scene": R_to_q(Rt_scene[:3, :3]), "t_scene": Rt_scene[:3, 3:4]}
