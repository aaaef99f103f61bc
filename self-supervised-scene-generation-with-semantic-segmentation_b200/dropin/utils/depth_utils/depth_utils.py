"""Import path of the reference module: `from utils.depth_utils.depth_utils import Depth2Normals` (reference train.py:22)."""
from spsg_b200.depth_utils import (Depth2Normals, bilateral_filter_floatmap, compute_normals,  # noqa: F401
                                   convert_depth_to_cameraspace, median_fill_depthmap)
