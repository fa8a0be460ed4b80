"""B200-native (sm_100a) cost-volume hot path of TransMVSNet behind the reference's function signatures."""
from .ops import (aggregate, cost_volume, depth_hypotheses, depth_regression, finalize_maps, depth_wta, fold_pixelwise_net, homo_warping,  # noqa: F401
                  pack_sources, pixelwise_aggregate, reference_arithmetic, softmax_wta)
from .depthnet import DepthNet, PixelwiseNet, patch_reference  # noqa: F401

__version__ = "0.2.0"
