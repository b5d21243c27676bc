"""Generate tests/golden/reference_layers.npz: ALL layers of the UNMODIFIED reference DeepFNet.forward
(deepFEPE/models/DeepFNet.py:429-554; depth 5, fp32, CPU) on seeded synthetic batches, for the live option
combinations: default (C2-shaped 8 x 1000 and the C1 planar case), if_quality with 1 and 2 quality channels, if_img_w
(weights_im), is_test=True.  Run once in the build container:  python tests/golden/make_golden_layers.py

Sign of the null vector: the signed residual X f of one layer is an input channel of the next layer's network
(DeepFNet.py:487) and the sign of f = V[:, -1] is whatever LAPACK returns.  As in make_golden_offsets.py, `torch.svd`
(third-party, not reference code) is wrapped so that the last right singular vector of an [N,9] matrix has its
largest-magnitude entry positive -- the convention of fepe_fit_fwd.  Any sign is a valid SVD; the reference's own
source is untouched.  Same seed + same construction order => the product model starts from identical parameters.
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
from make_golden_offsets import svd_canonical_null_vector, _svd  # noqa: E402

IMAGE = [376, 1241, 3]
# tag, seed, ctor kwargs, (B, N, data seed, planar)
CASES = [
    ("c2", 101, dict(if_quality=False), (8, 1000, 41, False)),
    ("c1", 102, dict(if_quality=False), (1, 100, 31, True)),
    ("q1", 103, dict(if_quality=True, quality_size=1), (2, 300, 43, False)),
    ("q2", 104, dict(if_quality=True, quality_size=2), (2, 300, 44, False)),
    ("imgw", 105, dict(if_quality=False, if_img_w=True), (2, 300, 45, False)),
    ("test", 106, dict(if_quality=False, is_test=True), (2, 300, 46, False)),
]


def case_inputs(tag, B, N, dseed, planar, kw):
    """The seeded inputs of one case (also used by the tests to rebuild them bit for bit)."""
    from fepe_b200 import synth
    d = synth.make_batch(B, N, dseed, planar=planar, outlier_frac=0.0 if planar else 0.3)
    g = torch.Generator().manual_seed(1000 + dseed)
    extra = {}
    if kw.get("if_quality"):
        extra["quality"] = torch.rand(B, N, kw["quality_size"], generator=g)
    if kw.get("if_img_w"):
        extra["weights_im"] = 0.5 + torch.rand(B, 1, N, generator=g)
    return d, extra


def main():
    MG.install_stubs()
    sys.path.insert(0, MG.ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from deepFEPE.models.DeepFNet import DeepFNet
    torch.set_num_threads(4)
    out = {}
    cuda_backup = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # DeepFNet.__init__ calls .cuda() (:356)
    torch.svd = svd_canonical_null_vector
    try:
        for tag, seed, kw, (B, N, dseed, planar) in CASES:
            torch.manual_seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                net = DeepFNet(depth=5, image_size=IMAGE, is_cuda=False, if_cpu_svd=False, **kw)
            d, extra = case_inputs(tag, B, N, dseed, planar, kw)
            batch = {"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]),
                     "matches_good_unique_nums": torch.from_numpy(d["matches_good_unique_nums"]),
                     "t_scene_scale": torch.ones(B, 1, 1), **extra}
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                o = net(batch)
            out[f"{tag}_F_layers"] = torch.stack(o["out_layers"]).numpy()
            out[f"{tag}_res_layers"] = torch.stack(o["residual_layers"]).numpy()
            out[f"{tag}_epi_layers"] = torch.stack(o["epi_res_layers"]).numpy()
            out[f"{tag}_w_layers"] = torch.stack(o["weights_layers"]).numpy()
            out[f"{tag}_logits_layers"] = torch.stack(o["logits_layers"]).numpy()
            out[f"{tag}_F_est"] = o["F_est"].numpy()
            # Yardstick: how far the reference's OWN fp32 result is from the fp64 evaluation of the same network on the
            # same inputs, per layer and pair (through oracle.deepf_forward, which reproduces the reference bit for bit in
            # fp32 -- tests/test_all_layers.py).  The depth-5 recursion amplifies rounding on ill-conditioned pairs (one
            # pair of the C2-shaped batch: 2e-4 at layer 3, 1e-3 at layer 4, while every other pair stays ~1e-6), so a
            # fixed 1e-4 bar is only meaningful where the reference itself is that well defined.
            from oracle import fepe_oracle as O
            q = kw.get("quality_size", 0) if kw.get("if_quality") else 0
            ni, nu = O.build_error_estimator(4 + q).double(), O.build_error_estimator(7 + q).double()
            ni.load_state_dict(net.input_weights.fw.state_dict())
            nu.load_state_dict(net.update_weights.fw.state_dict())
            dbl = lambda t: t.double() if t is not None else None
            torch.svd = _svd
            with torch.no_grad():
                o64 = O.deepf_forward(batch["matches_xy_ori"].double(), IMAGE, ni, nu, depth=5, quality=dbl(extra.get("quality")),
                                      weights_im=dbl(extra.get("weights_im")), canonical_sign=True)
            torch.svd = svd_canonical_null_vector
            out[f"{tag}_F_ref_vs_fp64"] = torch.stack([O.sign_aligned_rel_err(o64["out_layers"][l].float(), o["out_layers"][l])
                                                       for l in range(5)]).numpy()
            out[f"{tag}_logits_ref_vs_fp64"] = torch.stack([(o64["logits_layers"][l].float() - o["logits_layers"][l]).abs().amax((1, 2))
                                                            for l in range(5)]).numpy()
            out[f"{tag}_epi_ref_vs_fp64"] = torch.stack([(o64["epi_res_layers"][l].float() - o["epi_res_layers"][l]).abs().amax((1, 2))
                                                         for l in range(4)]).numpy()
            out[f"{tag}_res_ref_vs_fp64"] = torch.stack([(o64["residual_layers"][l].float() - o["residual_layers"][l]).abs().amax(1)
                                                         for l in range(5)]).numpy()
            print(tag, "done: F_est[0] =", o["F_est"][0].flatten()[:3].tolist())
    finally:
        torch.Tensor.cuda = cuda_backup
        torch.svd = _svd
    path = os.path.join(HERE, "reference_layers.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
