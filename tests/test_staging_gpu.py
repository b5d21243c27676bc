"""End-to-end host API (fepe_b200.staging.StagedStep): one pinned H2D copy, both kernels, one D2H copy;
results equal the direct device-resident calls."""
import numpy as np
import pytest
import torch

from fepe_b200 import ops, synth
from fepe_b200.staging import StagedStep

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def test_staged_step_equals_direct_calls():
    B, N = 32, 500
    d = synth.make_batch(B, N, seed=8, weight_mode="softmax")
    aff = ops.hw_affine(d["image_size"])
    st = StagedStep(B, N, d["pts1_virt"].shape[1], torch.device("cuda"))
    host = st.pack(d)
    stream = torch.cuda.Stream()
    st.run(stream, aff, host=host)
    stream.synchronize()
    F_h, pose_h = st.results()
    m, w = T(d["matches_xy_ori"]).cuda(), T(d["weights"]).cuda()
    F, res, epi, _ = ops.fit_forward(m, w, aff)
    pose = ops.pose_forward(F, T(d["Ks"]).cuda(), aff, T(d["q_cam"]).cuda(), T(d["t_cam"]).cuda(),
                            T(d["delta_Rtijs_4_4"]).cuda(), T(d["pts1_virt"]).cuda(), T(d["pts2_virt"]).cuda())
    torch.cuda.synchronize()
    assert torch.equal(F_h, F.cpu())
    assert torch.equal(pose_h, pose[0].cpu())
    assert torch.equal(st.d_res.cpu(), res.cpu()) and torch.equal(st.d_epi.cpu(), epi.cpu())
    assert st.in_bytes == 4 * (B * N * 5 + B * (9 + 4 + 3 + 16) + 2 * B * 100 * 3 + 0) or st.in_bytes >= 4 * B * N * 5


def test_captured_step_replays_with_new_host_data():
    """StagedStep.capture / replay: the graph re-reads the pinned buffer on every replay, so new host data gives new
    results, equal to the eager path."""
    B, N = 16, 400
    dev = torch.device("cuda")
    d0 = synth.make_batch(B, N, seed=21, weight_mode="softmax")
    d1 = synth.make_batch(B, N, seed=22, weight_mode="inlier")
    aff = ops.hw_affine(d0["image_size"])
    st = StagedStep(B, N, d0["pts1_virt"].shape[1], dev)
    stream = torch.cuda.Stream()
    st.pack(d0, out=st.h_in)
    st.capture(stream, aff)
    ref = StagedStep(B, N, d0["pts1_virt"].shape[1], dev)
    for d in (d1, d0, d1):
        st.pack(d, out=st.h_in)
        st.replay()
        stream.synchronize()
        F_g, pose_g = (t.clone() for t in st.results())
        ref.run(stream, aff, host=ref.pack(d))
        stream.synchronize()
        F_e, pose_e = ref.results()
        assert torch.equal(F_g, F_e) and torch.equal(pose_g, pose_e)
