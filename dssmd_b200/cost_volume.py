"""Drop-in for upstream models/UtilsNet/cost_volume.py (CostVolume, :5-53).

Deviation from upstream, on purpose: upstream's forward ends with a bare
`return` and therefore returns None for every similarity measure
(cost_volume.py:52-53); this module returns the volume.  Only 'correlation' is
on the hot path and runs the CUDA kernel.  'difference' and 'concat' are pure
copies that no upstream model instantiates; they are out of scope and raise.
"""
from __future__ import annotations

import torch.nn as nn

from .submoudles import build_corr


class CostVolume(nn.Module):
    def __init__(self, max_disp, feature_similarity='correlation'):
        """Construct cost volume based on different similarity measures.

        Args:
            max_disp: max disparity candidate
            feature_similarity: type of similarity measure
        """
        super().__init__()
        self.max_disp = max_disp
        self.feature_similarity = feature_similarity

    def forward(self, left_feature, right_feature):
        if self.feature_similarity == 'correlation':
            return build_corr(left_feature, right_feature, self.max_disp)  # [B, D, H, W]
        if self.feature_similarity in ('difference', 'concat'):
            raise NotImplementedError(
                f"CostVolume('{self.feature_similarity}') is outside the accelerated path "
                "(unused by every upstream model); use upstream's module for it")
        raise NotImplementedError  # same as upstream cost_volume.py:50
