"""leg_slam_b200 -- B200-native (sm_100a) implementation of LEG-SLAM's differentiable
Gaussian-splat mapping hot path behind the reference's rasterizer interface.

The compute lives in liblgs.so (csrc/, C ABI in include/lgs.h); this package is the
host-side mirror of the reference's Python/libtorch operator interface on top of it.
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
from .rasterize_points import mark_visible  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .query import cosine_image, cosine_query, heat_colors, heatmap_render, relevance_scores  # noqa: F401
from .renderer import GaussianModelView, GaussianPipelineParams, GaussianRenderer, KeyframeView  # noqa: F401
