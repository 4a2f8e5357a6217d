"""Swap the CUDA ops into an (unmodified) upstream DSSMD checkout.

    import sys; sys.path.insert(0, "/path/to/DSSMD")
    import dssmd_b200; dssmd_b200.install()
    from models.DSSMD import DSSMMD        # now calls the B200 kernels

Upstream's models obtain `build_corr` through `from models.UtilsNet.submoudles
import *` (models/DSSMD.py:9 and the three siblings), so the name must be rebound
both in the defining module and in every module that star-imported it.  The same
holds for disp_warp / LRwarp_error / CostVolume.
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
    # haze synthesis on the input side of the training step (trainfiles/dssmd_traner_new.py:22, 250-256)
    # attention along the epipolar line (models/BidNet/BidNet.py:6, 115)
    "models.BidNet.STTM_Simple": {"STTM_Single": sttm.STTM_Single},
    # ASM reconstruction loss (trainfiles/dssmd_traner_new.py:23, 175, 298-302)
    "losses.DSSMDLoss": {"RecoveredCleanImagesLoss": losses.RecoveredCleanImagesLoss,
                         "RecoveredCleanImagesLossV2": losses.RecoveredCleanImagesLossV2},
    "DeBug.inference": {"convert_disp_to_depth": haze.convert_disp_to_depth, "depth2trans": haze.depth2trans,
                        "recover_haze_images": haze.recover_haze_images},
}
# modules that copy those names into their own namespace at import time
_IMPORTERS = ["models.DSSMD", "models.DSSMD_new", "models.SSMDNet", "models.SDNet",
              "trainfiles.dssmd_traner_new", "trainfiles.dssmd_trainer", "trainfiles.dssmd_disp_only",
              "trainfiles.bidnet_trainer", "trainfiles.ffanet_trainer", "trainfiles.ffanet_trainer_kitti",
              "trainfiles.ffanet_trainer_ddp", "DeBug.test_model", "DeBug.test_model_only", "models.BidNet.BidNet"]

_saved = {}


def install(import_missing: bool = True):
    """Rebind the hot-path functions in the upstream modules.  Returns the list
    of (module, attribute) pairs that were patched.  With import_missing=True the
    defining modules are imported first (upstream must be on sys.path)."""
    patched = []
    replacements = {}
    for modname, attrs in _DEFINING.items():
        mod = sys.modules.get(modname)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(modname)
            except Exception:  # DeBug.inference needs skimage / matplotlib; patched only where it imports
                mod = None
        if mod is None:
            continue
        for attr, new in attrs.items():
            old = getattr(mod, attr, None)
            if old is None or old is new:
                continue
            _saved.setdefault((modname, attr), old)
            replacements[id(old)] = new
            setattr(mod, attr, new)
            patched.append((modname, attr))
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
