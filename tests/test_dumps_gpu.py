"""GPU check of the dump reader's batch path (SURVEY.md 8f rank 4 + 3): KittiCorrDump.batch() uploads the samples and
builds E, F, the virtual correspondences and q / t on the device (fepe_gt_virt); compared with what the reference's
dataset class computed per sample on the host for the same dump (tests/golden/dump_ref.npz), then fed to DeepFNet."""
import os

import numpy as np
import pytest
import torch

from fepe_b200 import dumps
from test_dumps_host import materialise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "dump_ref.npz"), allow_pickle=False))


def test_batch_matches_reference_dataset(ref, tmp_path):
    ds = materialise(ref, tmp_path, True)
    np.random.seed(int(ref["seed"]))
    b = ds.batch(range(len(ds)))
    torch.cuda.synchronize()
    host = lambda t: t.cpu().numpy()
    for k in ("K_ori", "K", "K_inv", "matches_all", "matches_good", "quality_good"):         # uploaded as read: bit exact
        np.testing.assert_array_equal(host(b[k]), ref["ref_" + k], err_msg=k)
    np.testing.assert_array_equal(host(b["relative_scene_poses"][1]), ref["ref_relative_scene_pose"])
    np.testing.assert_array_equal(b["matches_good_unique_nums"].numpy(), ref["ref_matches_good_unique_nums"])
    for k in ("E", "q_cam", "t_cam", "q_scene", "t_scene"):                                    # built on the device
        assert tuple(b[k].shape) == ref["ref_" + k].shape, k
        np.testing.assert_allclose(host(b[k]), ref["ref_" + k], rtol=0, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(host(b["F"]), ref["ref_F"], rtol=0, atol=5e-7)    # the reference's float32 cancellation
    for k in ("pts1_virt", "pts2_virt"):
        assert np.abs(host(b[k]) - ref["ref_" + k]).max() <= 1e-2, k     # pixels; F built in fp64 here, float32 there
    for k in ("pts1_virt_normalized", "pts2_virt_normalized"):
        assert np.abs(host(b[k]) - ref["ref_" + k]).max() <= 2e-5, k


def test_data_batch_feeds_the_model(ref, tmp_path):
    from fepe_b200.models import DeepFNet
    ds = materialise(ref, tmp_path, False)
    np.random.seed(0)
    b = ds.batch([0, 1, 2])
    db = dumps.KittiCorrDump.data_batch(b, if_quality=True)
    assert db["matches_xy_ori"].shape == (3, 1000, 4) and db["quality"].shape == (3, 1000, 2)
    # matches_xy = K^-1 applied to both points (Train_model_pipeline.py:401-413)
    m, Ki = db["matches_xy_ori"].double(), db["K_invs"].double()
    x1 = torch.cat((m[:, :, :2], torch.ones_like(m[:, :, :1])), 2) @ Ki.transpose(1, 2)
    np.testing.assert_allclose(db["x1_normalizedK"].cpu().numpy(), (x1[:, :, :2] / x1[:, :, 2:]).cpu().numpy(), atol=1e-5)
    torch.manual_seed(0)
    net = DeepFNet(depth=3, image_size=ds.sizerHW, if_quality=True, quality_size=2).cuda()
    with torch.no_grad():
        outs = net(db)
    assert outs["F_est"].shape == (3, 3, 3) and bool(torch.isfinite(outs["F_est"]).all())
    assert len(outs["out_layers"]) == 3 and outs["logits"].shape == (3, 1000)
