"""ErrorEstimator -- the per-correspondence weight network (deepFEPE/models/ErrorEstimators.py:14-68).

Shared MLP over the N correspondences: five blocks of (1x1 Conv1d -> InstanceNorm1d(affine) ->
LeakyReLU(0.01)) with 64/128/1024/512/256 channels and a final 1x1 Conv1d.  The nn.Sequential is laid
out exactly like the reference's live branch (if_bn=False, :46-64) so checkpoints interchange:
keys fw.{0,3,6,9,12,15}.{weight,bias} (convs) and fw.{1,4,7,10,13}.{weight,bias} (norm affine).

The layers run through PyTorch's cuDNN/cuBLAS kernels in this round (library path); the tcgen05
GEMM with fused InstanceNorm statistics is row "next" in DESIGN.md.
"""
import torch
import torch.nn as nn


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
        self.tensor_cores = False       # opt-in: bf16 tcgen05 path under torch.no_grad() (fepe_b200/mlp_tc.py)
        self.tensor_cores_training = False   # opt-in: tcgen05 forward AND backward under autograd
        self._tc = None
        self.last_softmax = None        # softmax over N of the last tensor-core evaluation (fused in its last kernel)

    def forward(self, data):
        if (self.tensor_cores and not torch.is_grad_enabled() and data.is_cuda and self.fw[-1].out_channels == 1
                and data.shape[1] <= 8):
            if self._tc is None:
                from ..mlp_tc import TensorCoreMLP
                self._tc = TensorCoreMLP(self.fw)
            logits, self.last_softmax = self._tc(data.float())
            return logits
        self.last_softmax = None
        if (self.tensor_cores_training and torch.is_grad_enabled() and data.is_cuda and self.fw[-1].out_channels == 1
                and data.shape[1] <= 8):
            from ..mlp_tc import TensorCoreMLPFunction, module_params
            return TensorCoreMLPFunction.apply(data.float(), *module_params(self.fw))
        # the reference computes these 1x1 convolutions in fp32 (torch 1.3 had no TF32); keep cuDNN from
        # silently dropping to TF32 so that weights / logits match it to fp32 accuracy
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self.fw(data)
