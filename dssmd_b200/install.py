"""Swap the CUDA ops into an (unmodified) upstream DSSMD checkout.

    import sys; sys.path.insert(0, "/path/to/DSSMD")
    import dssmd_b200; dssmd_b200.install()
    from models.DSSMD import DSSMMD        # now calls the B200 kernels

Upstream's models obtain `build_corr` through `from models.UtilsNet.submoudles
import *` (models/DSSMD.py:9 and the three siblings), so the name must be rebound
both in the defining module and in every module that star-imported it.  The same
holds for disp_warp / LRwarp_error / CostVolume.

Two kinds of patch:

* the hot path (build_corr, CostVolume, disp_warp, LRwarp_error) is REPLACED: every
  dtype / shape upstream accepts is handled by a kernel or raises, and a CPU tensor
  raises (no CPU fallback by design);
* the kernels either side of it (`aux=True`: haze synthesis, the ASM reconstruction
  loss, the epipolar attention of STTM_Single) have narrower limits than upstream's
  torch expressions (fp32, aligned widths, no gradient into the images ...).  They are
  GUARDED: upstream's classes keep their identity, constructor and parameters and only
  their `forward` is patched; upstream's functions are wrapped.  Inputs inside a
  kernel's limits take the kernel, everything else runs upstream's own saved code, so a
  model that worked before install() still works after it.
"""
from __future__ import annotations

import importlib
import sys

from . import cost_volume, disparity_warper, haze, losses, sttm, submoudles

# upstream module -> {attribute: replacement}
_DEFINING = {
    "models.UtilsNet.submoudles": {"build_corr": submoudles.build_corr},
    "models.UtilsNet.cost_volume": {"CostVolume": cost_volume.CostVolume},
    "models.UtilsNet.stereo_op": {"disp_warp": disparity_warper.disp_warp,
                                  "LRwarp_error": disparity_warper.LRwarp_error},
    "utils.disparity_warper": {"disp_warp": disparity_warper.disp_warp,
                               "LRwarp_error": disparity_warper.LRwarp_error},
}
# guarded function wrappers: upstream module -> {attribute: kernel function}
# haze synthesis on the input side of the training step (trainfiles/dssmd_traner_new.py:22, 250-256)
_GUARDED_FUNCTIONS = {
    "DeBug.inference": {"convert_disp_to_depth": haze.convert_disp_to_depth, "depth2trans": haze.depth2trans,
                        "recover_haze_images": haze.recover_haze_images},
}
# guarded `forward` methods: upstream module -> {class name: factory(upstream_forward) -> forward}
_GUARDED_CLASSES = {
    # attention along the epipolar line (models/BidNet/BidNet.py:6, 115)
    "models.BidNet.STTM_Simple": {"STTM_Single": sttm.guarded_forward},
    # ASM reconstruction loss (trainfiles/dssmd_traner_new.py:23, 175, 298-302)
    "losses.DSSMDLoss": {"RecoveredCleanImagesLoss": losses.guarded_asm_forward,
                         "RecoveredCleanImagesLossV2": losses.guarded_asm_forward},
}
# modules that copy those names into their own namespace at import time
_IMPORTERS = ["models.DSSMD", "models.DSSMD_new", "models.SSMDNet", "models.SDNet",
              "trainfiles.dssmd_traner_new", "trainfiles.dssmd_trainer", "trainfiles.dssmd_disp_only",
              "trainfiles.bidnet_trainer", "trainfiles.ffanet_trainer", "trainfiles.ffanet_trainer_kitti",
              "trainfiles.ffanet_trainer_ddp", "DeBug.test_model", "DeBug.test_model_only", "models.BidNet.BidNet"]

_saved = {}          # (module name, attribute) -> upstream object
_saved_methods = []  # (class, method name, upstream function)


def _module(modname, import_missing):
    mod = sys.modules.get(modname)
    if mod is None and import_missing:
        try:
            mod = importlib.import_module(modname)
        except Exception:  # DeBug.inference needs skimage / matplotlib; patched only where it imports
            mod = None
    return mod


def install(import_missing: bool = True, aux: bool = True):
    """Rebind the hot-path functions in the upstream modules.  Returns the list
    of (module, attribute) pairs that were patched.  With import_missing=True the
    defining modules are imported first (upstream must be on sys.path).  aux=False
    leaves the haze functions, the ASM loss and STTM_Single alone."""
    patched = []
    replacements = {}

    def rebind(mod, modname, attr, old, new):
        _saved.setdefault((modname, attr), old)
        replacements[id(old)] = new
        setattr(mod, attr, new)
        patched.append((modname, attr))

    for modname, attrs in _DEFINING.items():
        mod = _module(modname, import_missing)
        if mod is None:
            continue
        for attr, new in attrs.items():
            old = getattr(mod, attr, None)
            if old is None or old is new:
                continue
            rebind(mod, modname, attr, old, new)
    if aux:
        for modname, attrs in _GUARDED_FUNCTIONS.items():
            mod = _module(modname, import_missing)
            if mod is None:
                continue
            for attr, kernel_fn in attrs.items():
                old = getattr(mod, attr, None)
                if old is None or getattr(old, "__kernel__", None) is kernel_fn:
                    continue
                rebind(mod, modname, attr, old, haze.guarded(kernel_fn, old))
        for modname, classes in _GUARDED_CLASSES.items():
            mod = _module(modname, import_missing)
            if mod is None:
                continue
            for cname, factory in classes.items():
                cls = getattr(mod, cname, None)
                fwd = None if cls is None else cls.__dict__.get("forward")
                if fwd is None or hasattr(fwd, "__wrapped_upstream__"):
                    continue
                _saved_methods.append((cls, "forward", fwd))
                setattr(cls, "forward", factory(fwd))   # in place: every importer of the class sees it
                patched.append((modname, cname + ".forward"))
    for modname in _IMPORTERS:
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for attr, val in list(vars(mod).items()):
            new = replacements.get(id(val))
            if new is not None:
                _saved.setdefault((modname, attr), val)
                setattr(mod, attr, new)
                patched.append((modname, attr))
    return patched


def uninstall():
    """Undo install()."""
    for (modname, attr), old in list(_saved.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, attr, old)
    _saved.clear()
    for cls, name, fn in _saved_methods:
        setattr(cls, name, fn)
    del _saved_methods[:]
