"""TEST INFRASTRUCTURE ONLY: locate and import the UNMODIFIED reference (eric-yyjau/pytorch-deepFEPE @ 7f3e775).

The reference is pure Python.  `oracle/make_ref.py` copies the handful of files this path needs -- byte for byte,
nothing edited -- from /root/reference into the git-ignored directory `oracle/_ref/` (it travels to the GPU box with
the snapshot the way a built .so does; the repo history never contains reference sources).  This module puts that
copy (or /root/reference itself when it is mounted) on sys.path, installs import stubs for the packages the
reference imports at module scope but never touches on this path, and hands the modules out.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference) may import this.
"""
from __future__ import annotations

import contextlib
import io
import logging
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_MOUNT = "/root/reference"

# files of the reference that the hot path, its callers (loss glue, model loader) and their module-scope imports need
REF_FILES = (
    "deepFEPE/models/__init__.py", "deepFEPE/models/DeepFNet.py", "deepFEPE/models/ErrorEstimators.py",
    "deepFEPE/models/GoodCorresNet.py", "deepFEPE/models/model_utils.py",
    "deepFEPE/dsac_tools/__init__.py", "deepFEPE/dsac_tools/utils_F.py", "deepFEPE/dsac_tools/utils_geo.py",
    "deepFEPE/dsac_tools/utils_misc.py", "deepFEPE/dsac_tools/utils_vis.py", "deepFEPE/dsac_tools/utils_opencv.py",
    "deepFEPE/train_good_utils.py", "deepFEPE/settings.py",
    "deepFEPE/utils/__init__.py", "deepFEPE/utils/loader.py",
)


def reference_root():
    """Directory holding `deepFEPE/` of the unmodified reference, or None."""
    for root in (REF_COPY, REF_MOUNT):
        if os.path.exists(os.path.join(root, "deepFEPE", "models", "DeepFNet.py")):
            return root
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Packages the reference imports at module scope and never uses on this path: matplotlib
    (utils_vis.py:2), pebble / superpoint.* (train_good_utils.py:24,40-47), tensorboardX / imgaug (utils/loader.py:20,31)."""
    noop = lambda *a, **k: None
    for name, attrs in (
        ("matplotlib", dict(use=noop)), ("matplotlib.pyplot", {}), ("matplotlib.cm", {}), ("mpl_toolkits", {}),
        ("mpl_toolkits.mplot3d", dict(Axes3D=object)), ("pebble", dict(ProcessPool=object)),
        ("superpoint", {}), ("superpoint.utils", {}),
        ("superpoint.utils.logging", dict(logging=logging, toRed=str, toCyan=str)),
        ("superpoint.utils.utils", dict(tensor2array=noop, save_checkpoint=noop, load_checkpoint=noop,
                                        save_path_formatter=noop, flattenDetection=noop)),
        ("superpoint.utils.var_dim", dict(toNumpy=noop, squeezeToNumpy=noop)),
        ("tensorboardX", dict(SummaryWriter=object)), ("imgaug", dict(augmenters=types.ModuleType("augmenters"))),
        ("imgaug.augmenters", {}),
    ):
        if name not in sys.modules:
            try:
                __import__(name)
                continue
            except Exception:
                pass
            _stub(name, **attrs)


_cached = None


def import_reference():
    """-> namespace with the reference's own modules / classes:
    Fit, NormalizeAndExpand_HW, DeepFNet, ErrorEstimator, utils_F, utils_geo, utils_misc, tgu (train_good_utils),
    root.  Raises RuntimeError when neither oracle/_ref nor /root/reference exists."""
    global _cached
    if _cached is not None:
        return _cached
    root = reference_root()
    if root is None:
        raise RuntimeError("the unmodified reference is not available: run `python oracle/make_ref.py` where "
                           "/root/reference is mounted (it fills the git-ignored oracle/_ref/)")
    for p in (os.path.join(root, "deepFEPE"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    install_stubs()
    with contextlib.redirect_stdout(io.StringIO()):
        from deepFEPE.models.DeepFNet import Fit, NormalizeAndExpand_HW, DeepFNet
        from deepFEPE.models.ErrorEstimators import ErrorEstimator
        from deepFEPE.dsac_tools import utils_F, utils_geo, utils_misc
        import train_good_utils as tgu
    _cached = types.SimpleNamespace(Fit=Fit, NormalizeAndExpand_HW=NormalizeAndExpand_HW, DeepFNet=DeepFNet,
                                    ErrorEstimator=ErrorEstimator, utils_F=utils_F, utils_geo=utils_geo,
                                    utils_misc=utils_misc, tgu=tgu, root=root)
    return _cached


@contextlib.contextmanager
def quiet():
    """The reference prints from inside its geometry helpers (utils_misc._homo, dsac_tools/utils_misc.py:60)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
