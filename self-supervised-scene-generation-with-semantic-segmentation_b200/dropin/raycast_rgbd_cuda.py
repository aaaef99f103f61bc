"""Top-level module name of the reference's native extension (raycast_rgbd.py:7 `import raycast_rgbd_cuda`)."""
from spsg_b200.raycast_rgbd_cuda import backward, construct_dense_sparse_mapping, forward, raycast_occ  # noqa: F401
