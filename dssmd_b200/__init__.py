"""dssmd_b200 -- B200-native correlation cost volume + disparity warp for DSSMD.

Host-side mirror of the upstream interface (same names, arguments, return
values and error behaviour) over a C-ABI library of hand-written sm_100a CUDA
kernels.  No CPU fallback, no Triton, no multi-backend dispatch.
"""
from .cost_volume import CostVolume
from .disparity_warper import LRwarp_error, disp_warp, meshgrid, normalize_coords, read_pfm
from .functional import check_negative_disparity_now, resize_bilinear, set_check_negative_disparity
from .haze import convert_disp_to_depth, depth2trans, recover_haze_images, synthesize_haze
from .install import install, uninstall
from .losses import RecoveredCleanImagesLoss, RecoveredCleanImagesLossV2, photometric_l1, photometric_loss, recovered_clean_l1
from .stereo_op import print_tensor_shape
from .sttm import STTM_Single, epipolar_attention
from .submoudles import build_corr, build_corr_cat

__all__ = [
    "build_corr", "build_corr_cat", "CostVolume", "disp_warp", "LRwarp_error", "normalize_coords", "meshgrid",
    "read_pfm", "print_tensor_shape", "resize_bilinear", "set_check_negative_disparity", "check_negative_disparity_now",
    "install", "uninstall", "photometric_l1", "photometric_loss", "RecoveredCleanImagesLoss", "RecoveredCleanImagesLossV2", "recovered_clean_l1",
    "STTM_Single", "epipolar_attention", "convert_disp_to_depth", "depth2trans", "recover_haze_images", "synthesize_haze",
]
__version__ = "0.1.0"
