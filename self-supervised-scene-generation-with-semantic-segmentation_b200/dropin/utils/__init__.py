# Namespace-style package so that the reference's other `utils.*` sub-packages (depth_utils, color_utils_cpu,
# marching_cubes) found later on sys.path keep importing next to this drop-in `utils.raycast_rgbd`.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
