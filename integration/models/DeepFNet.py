"""What a maintainer puts in place of deepFEPE/models/DeepFNet.py: the reference's
`utils/loader.py:117-129` resolves model.name "GoodCorresNet_layers_deepF" with `from models.DeepFNet import DeepFNet`,
so re-exporting the B200 classes under that module name is the whole integration.  tests/test_reference_callers.py
runs the reference's own modelLoader, get_all_loss_DeepF and get_Rt_loss against this file."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "pytorch-deepfepe_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from fepe_b200.models.DeepFNet import DeepFNet, Fit, NormalizeAndExpand_HW  # noqa: E402,F401
