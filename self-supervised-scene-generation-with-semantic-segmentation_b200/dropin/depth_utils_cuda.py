"""Top-level module name of the reference's second native extension (depth_utils.py:8 `import depth_utils_cuda`)."""
from spsg_b200.depth_utils_cuda import (bilateral_filter_floatmap, compute_normals,  # noqa: F401
                                        convert_depth_to_cameraspace, median_fill_depthmap)
