"""ErrorEstimator -- the per-correspondence weight network (deepFEPE/models/ErrorEstimators.py:14-68).

Shared MLP over the N correspondences: five blocks of (1x1 Conv1d -> InstanceNorm1d(affine) ->
LeakyReLU(0.01)) with 64/128/1024/512/256 channels and a final 1x1 Conv1d.  The nn.Sequential is laid
out exactly like the reference's live branch (if_bn=False, :46-64) so checkpoints interchange:
keys fw.{0,3,6,9,12,15}.{weight,bias} (convs) and fw.{1,4,7,10,13}.{weight,bias} (norm affine).

Arithmetic (``set_path``):
  "tc32"  (default) hand-written tcgen05 GEMMs at the reference's fp32 accuracy: operands split into fp16 (hi, lo)
          pairs, three MMAs per product, fp32 accumulation, fp64 InstanceNorm statistics (csrc/fepe_mlp32.cu,
          fepe_b200/mlp32.py); inference and training (forward + backward), output_size 1 and 4.
  "bf16"  the faster bf16 tcgen05 path (csrc/fepe_mlp.cu, fepe_b200/mlp_tc.py): opt-in, logits within ~3e-2 of fp32.
  "torch" PyTorch's fp32 library kernels: A/B baseline only, never the default.
CPU tensors raise: there is no CPU path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

PATHS = ("tc32", "bf16", "torch")


class ErrorEstimator(nn.Module):
    def __init__(self, input_size, output_size=1, if_bn=False):
        super().__init__()
        if if_bn:
            raise NotImplementedError("ErrorEstimator(if_bn=True) is never constructed by the reference's "
                                      "DeepFNet (DeepFNet.py:339-342) and is not provided")
        chans = [input_size, 64, 128, 1024, 512, 256]
        layers = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            layers += [nn.Conv1d(cin, cout, kernel_size=1, bias=True),
                       nn.InstanceNorm1d(cout, affine=True),
                       nn.LeakyReLU(inplace=True)]
        layers.append(nn.Conv1d(256, output_size, kernel_size=1, bias=True))
        self.fw = nn.Sequential(*layers)
        self.input_size, self.output_size = input_size, output_size
        self.path = "tc32"
        self._tc32 = None
        self._tc = None
        self.last_softmax = None        # softmax over N of the last evaluation when the kernels produced it

    # -- configuration ---------------------------------------------------------------------------
    def set_path(self, path: str):
        if path not in PATHS:
            raise ValueError(f"ErrorEstimator.set_path: {path!r} is not one of {PATHS}")
        if path == "bf16" and (self.output_size != 1 or self.input_size > 8):
            path = "tc32"               # the bf16 kernels cover the one-logit networks with <= 8 inputs only
        self.path = path
        return self

    # legacy switches of round 1 (bf16 opt-in under no_grad / under autograd)
    @property
    def tensor_cores(self):
        return self.path == "bf16"

    @tensor_cores.setter
    def tensor_cores(self, on):
        self.set_path("bf16" if on else "tc32")

    @property
    def tensor_cores_training(self):
        return self.path == "bf16"

    @tensor_cores_training.setter
    def tensor_cores_training(self, on):
        self.set_path("bf16" if on else "tc32")

    # -- evaluation ------------------------------------------------------------------------------
    def forward(self, data):
        """data [B, Cin, N] -> logits [B, out, N] (the reference's interface, ErrorEstimators.py:66-68)."""
        B, _, N = data.shape
        return self.forward_parts(None, None, [data.permute(0, 2, 1)], B, N)

    def forward_parts(self, matches, affine, extras, B, N):
        """The same network evaluated straight from the model's inputs: `matches` [B,N,4] pixels with the
        NormalizeAndExpand_HW affine (channels ((a x + b) + 1) / 2, DeepFNet.get_input :377-389) followed by the channel
        groups `extras` ([B,N,c] or [B,N]) in the order of the reference's torch.cat (:387-389, :487).  On the kernel
        paths nothing is concatenated or permuted; `matches` may be None (all channels in `extras`)."""
        ref = matches if matches is not None else extras[0]
        if not ref.is_cuda:
            raise RuntimeError("fepe_b200.ErrorEstimator needs CUDA tensors: there is no CPU path")
        self.last_softmax = None
        # (m.weight / m.bias, not fw.parameters(): nn.DataParallel replicas carry no registered parameters)
        grad = torch.is_grad_enabled() and (any(t.requires_grad for m in self.fw if hasattr(m, "weight")
                                                for t in (m.weight, m.bias))
                                            or any(t.requires_grad for t in extras)
                                            or (matches is not None and matches.requires_grad))
        if self.path == "tc32":
            from .. import mlp32
            if not grad:
                # (nn.DataParallel replicas are shallow copies: a cached evaluator must belong to THIS replica's layers)
                if self._tc32 is None or self._tc32.fw is not self.fw:
                    self._tc32 = mlp32.MLP32(self.fw)
                logits, self.last_softmax = self._tc32(matches, affine, [t.float() for t in extras], B, N)
                return logits
            if hasattr(mlp32, "MLP32Function"):
                return mlp32.mlp32_autograd(self.fw, matches, affine, extras, B, N)
        data = self._features(matches, affine, extras)
        if self.path == "bf16" and self.output_size == 1 and data.shape[1] <= 8:
            from ..mlp_tc import TensorCoreMLP, TensorCoreMLPFunction, module_params
            if not grad:
                if self._tc is None or self._tc.fw is not self.fw:
                    self._tc = TensorCoreMLP(self.fw)
                logits, self.last_softmax = self._tc(data.float().contiguous())
                return logits
            return TensorCoreMLPFunction.apply(data.float(), *module_params(self.fw))
        # library fp32 (the reference's arithmetic: torch 1.3 had no TF32) -- keep cuDNN from dropping to TF32
        cd = torch.backends.cudnn
        with cd.flags(enabled=cd.enabled, benchmark=cd.benchmark, deterministic=cd.deterministic, allow_tf32=False):
            return self.fw(data)

    @staticmethod
    def _features(matches, affine, extras):
        """[B, Cin, N] feature tensor with torch ops (library / bf16 paths, and tc32 under autograd until its backward
        kernels take over): the reference's get_input + cat."""
        feats = []
        if matches is not None:
            ax, bx, ay, by = affine
            feats += [((matches[:, :, 0:1] * ax + bx) + 1) / 2, ((matches[:, :, 1:2] * ay + by) + 1) / 2,
                      ((matches[:, :, 2:3] * ax + bx) + 1) / 2, ((matches[:, :, 3:4] * ay + by) + 1) / 2]
        feats += [e.unsqueeze(2) if e.dim() == 2 else e for e in extras]
        return torch.cat(feats, 2).permute(0, 2, 1)
