"""Import path of the reference wrapper: `from utils.raycast_rgbd.raycast_rgbd import RaycastRGBD, RaycastOcc`
(reference train.py:20-21, test_scene.py:17)."""
from spsg_b200.raycast_rgbd import RayCastRGBDFunction, RaycastOcc, RaycastRGBD  # noqa: F401
