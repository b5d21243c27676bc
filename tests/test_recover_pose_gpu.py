"""GPU parity of fepe_recover_pose (validation pose recovery) through the C ABI against the numpy
oracle (pinned to cv2.recoverPose and to the reference's goodCorr_eval_nondecompose, tests/test_recover_pose_host.py)
and against the committed outputs of the reference function itself."""
import os

import numpy as np
import pytest
import torch

from fepe_b200 import ops, synth
from fepe_b200.validation import goodCorr_eval_nondecompose_batch
from oracle import recover_pose_oracle as RO

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def test_against_reference_function_outputs():
    ref = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "recover_pose_ref.npz"), allow_pickle=False))
    n = ref["E"].shape[0]
    out, mask = ops.recover_pose(T(ref["E"]).cuda(), T(ref["K"]).cuda(), T(ref["matches"]).cuda(), T(ref["Rt"]).cuda(),
                                 want_mask=True)
    out, mask = out[0].cpu().numpy(), mask[0].cpu().numpy()
    for i in range(n):
        M = np.hstack((out[i, :9].reshape(3, 3), out[i, 9:12].reshape(3, 1)))
        np.testing.assert_allclose(M, ref["M"][i], atol=2e-6)
        assert int(out[i, 12]) == int(ref["good"][i])
        np.testing.assert_array_equal(mask[i] > 0, ref["mask"][i] > 0)
        assert abs(out[i, 18] - ref["err"][i][0]) < 1e-3 and abs(out[i, 19] - ref["err"][i][1]) < 1e-3     # degrees


@pytest.mark.parametrize("B,N", [(8, 1000), (5, 333), (3, 2000)])
def test_against_oracle_layers_and_ragged(B, N):
    """Two 'layers' of essential matrices (exact and perturbed), ragged valid counts."""
    d = synth.make_batch(B, N, seed=500 + N)
    rng = np.random.default_rng(N)
    E = np.stack((d["E_gt"], d["E_gt"] + 0.04 * rng.normal(size=d["E_gt"].shape))).astype(np.float32)   # [2,B,3,3]
    nv = rng.integers(N // 2, N + 1, size=B).astype(np.int32)
    nv[0] = N
    res = goodCorr_eval_nondecompose_batch(T(d["matches_xy_ori"]).cuda(), T(E).cuda(), T(d["delta_Rtijs_4_4"]).cuda(),
                                           T(d["Ks"]).cuda(), n_valid=T(nv).cuda(), want_mask=True)
    torch.cuda.synchronize()
    for l in range(2):
        for b in range(B):
            K = d["Ks"][b].astype(np.float64)
            m = d["matches_xy_ori"][b][:nv[b]]
            good, R, t, mo, cnt = RO.recover_pose(E[l, b].astype(np.float64), m[:, :2], m[:, 2:], K[0, 0], (K[0, 2], K[1, 2]))
            M = res["M"][l, b].cpu().numpy()
            np.testing.assert_allclose(M, np.hstack((R, t[:, None])), atol=2e-6)
            assert int(res["num_inlier"][l, b]) == good
            assert sorted(res["counts"][l, b].tolist()) == sorted(cnt.tolist())
            mk = res["mask"][l, b].cpu().numpy()
            np.testing.assert_array_equal(mk[:nv[b]] > 0, mo)
            assert not mk[nv[b]:].any()
            eq, et = RO.pose_errors_vs_gt(R, t, d["delta_Rtijs_4_4"][b])
            assert abs(float(res["err_q"][l, b]) - eq) < 1e-3 and abs(float(res["err_t"][l, b]) - et) < 1e-3


def test_too_few_points_and_bad_arguments():
    d = synth.make_batch(2, 64, seed=9)
    nv = torch.tensor([3, 64], dtype=torch.int32).cuda()
    out, _ = ops.recover_pose(T(d["E_gt"]).cuda(), T(d["Ks"]).cuda(), T(d["matches_xy_ori"]).cuda(),
                              T(d["delta_Rtijs_4_4"]).cuda(), n_valid=nv)
    o = out[0, 0].cpu().numpy()
    np.testing.assert_allclose(o[:9].reshape(3, 3), np.eye(3))         # utils_F.py:948-952
    assert o[18] == 180.0 and o[19] == 90.0 and not o[9:12].any()
    with pytest.raises(RuntimeError):
        ops.recover_pose(T(d["E_gt"]), T(d["Ks"]).cuda(), T(d["matches_xy_ori"]).cuda())      # CPU tensor: no fallback
