"""Repo-root shim so that ``import util`` (the reference's module name, util.py) resolves to the B200 implementation of
the functions on the hot path: warping, Pano2PointCloud, apply_mask (relativepose_b200/util.py)."""
from relativepose_b200.util import *          # noqa: F401,F403
from relativepose_b200.util import warping, Pano2PointCloud, apply_mask, warping_device, pano2pointcloud_device, blend_completion_device  # noqa: F401
