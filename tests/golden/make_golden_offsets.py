"""Generate tests/golden/reference_offsets.npz: the UNMODIFIED reference DeepFNet with if_learn_offsets=True
(deepFEPE/models/DeepFNet.py:341-342, 369-373, 489-505) run in fp64 on the CPU of the build container, forward
and backward, on a seeded synthetic batch.  Pins the coordinate-gradient path (fepe_fit_bwd_coords) end to end
through the reference's own module.  Run once:  python tests/golden/make_golden_offsets.py

Sign of the null vector: the reference feeds the SIGNED residual X f of one layer to the networks of the next
(DeepFNet.py:487), and the sign of f = V[:, -1] is whatever LAPACK returns.  To make the multi-layer run comparable,
`torch.svd` (third-party, not reference code) is wrapped here so that the last right singular vector of an [N,9]
matrix has its largest-magnitude entry positive -- the convention of fepe_fit_fwd.  Any sign is a valid SVD; the
reference's own source is untouched.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (path setup + import stubs)

SEED, DEPTH, B, N = 78, 3, 2, 200

_svd = torch.svd


def svd_canonical_null_vector(A, *args, **kwargs):
    U, S, V = _svd(A, *args, **kwargs)
    if A.dim() == 2 and A.shape[1] == 9:
        v = V[:, -1].detach()
        flip = torch.ones(9, dtype=V.dtype)
        flip[-1] = torch.sign(v[v.abs().argmax()])
        U, V = U * flip, V * flip          # still A = U diag(S) V^T
    return U, S, V


def main():
    MG.install_stubs()
    from fepe_b200 import synth
    with contextlib.redirect_stdout(io.StringIO()):
        from deepFEPE.models.DeepFNet import DeepFNet
        from deepFEPE.dsac_tools import utils_F
    torch.set_num_threads(4)
    out = {}
    cuda_backup = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # DeepFNet.__init__ calls .cuda() (:356)
    torch.svd = svd_canonical_null_vector
    try:
        torch.manual_seed(SEED)
        with contextlib.redirect_stdout(io.StringIO()):
            net = DeepFNet(depth=DEPTH, image_size=[376, 1241, 3], if_quality=False, is_cuda=False, if_cpu_svd=False,
                           if_learn_offsets=True)
        out["state_keys"] = np.array(list(net.state_dict().keys()))
        net = net.double()
        # .double() does not reach the plain-tensor attributes of Fit / NormalizeAndExpand_HW (DeepFNet.py:98,128-133)
        for mod, names in ((net.fit, ("ones_b", "zero_b", "T_b", "mask")), (net.norm_HW, ("ones_b",))):
            for nm in names:
                setattr(mod, nm, getattr(mod, nm).double())
        d = synth.make_batch(B, N, 33, weight_mode="softmax", outlier_frac=0.1)
        batch = {"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]).double(),
                 "matches_good_unique_nums": torch.from_numpy(d["matches_good_unique_nums"]),
                 "t_scene_scale": torch.ones(B, 1, 1, dtype=torch.float64)}
        with contextlib.redirect_stdout(io.StringIO()):
            o = net(batch)
            # a sign-invariant loss touching every output the offsets influence
            loss = 0.0
            for Fl in o["out_layers"]:
                loss = loss + utils_F.compute_epi_residual(o["pts1"], o["pts2"], Fl, 0.1).mean()
            for r in o["residual_layers"]:
                loss = loss + 1e3 * (r ** 2).sum(1).mean()
            loss.backward()
        out["matches"] = d["matches_xy_ori"]
        out["offsets"] = o["offsets"].detach().numpy()
        out["F_layers"] = torch.stack(o["out_layers"]).detach().numpy()
        out["w_layers"] = torch.stack(o["weights_layers"]).detach().numpy()
        out["epi_layers"] = torch.stack(o["epi_res_layers"]).detach().numpy()
        out["pts1"] = o["pts1"].detach().numpy()
        out["loss"] = np.array(float(loss))
        # gradients of the small parameter tensors of each net (the big ones would bloat the fixture)
        for name, p in net.named_parameters():
            if p.grad is not None and p.numel() <= 4096:
                out["grad/" + name] = p.grad.numpy()
    finally:
        torch.Tensor.cuda = cuda_backup
        torch.svd = _svd
    path = os.path.join(HERE, "reference_offsets.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, loss {float(loss):.6f}")


if __name__ == "__main__":
    main()
