"""Host-side mirror of the reference's model files (deepFEPE/models/*.py): same class names, constructor
signatures, forward I/O and state_dict keys, with the hot path running in libfepe_b200.so."""
from .DeepFNet import DeepFNet, Fit, NormalizeAndExpand_HW  # noqa: F401
from .ErrorEstimators import ErrorEstimator  # noqa: F401
from .GoodCorresNet import GoodCorresNet  # noqa: F401
