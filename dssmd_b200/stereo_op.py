"""Drop-in for upstream models/UtilsNet/stereo_op.py (the module the models import)."""
from .disparity_warper import LRwarp_error, disp_warp, meshgrid, normalize_coords  # noqa: F401

__all__ = ["normalize_coords", "meshgrid", "disp_warp", "LRwarp_error", "print_tensor_shape"]


def print_tensor_shape(inputs):
    """Print the shape of a tensor or of every tensor in a list/tuple (upstream stereo_op.py:75-82)."""
    if isinstance(inputs, (list, tuple)):
        for value in inputs:
            print(value.shape)
    else:
        print(inputs.shape)
