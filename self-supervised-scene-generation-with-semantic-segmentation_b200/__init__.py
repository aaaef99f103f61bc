"""B200-native SPSG-semantic raycaster: hot path only.

Sub-modules (import as ``spsg_b200.<name>``):
  _native            ctypes binding of the C ABI in include/spsg_raycast.h (lib/libspsg_raycast.so)
  raycast_rgbd_cuda  drop-in for the reference's native extension module of the same name
  raycast_rgbd       drop-in for torch/utils/raycast_rgbd/raycast_rgbd.py (RaycastRGBD, RaycastOcc)
  losses             2D depth / colour / semantic losses consuming the renderings (fused + literal)
  synthetic          seeded synthetic chunks, cameras and frames
  parallel           one-process-per-GPU sharding helpers (chunk x view batches, NCCL grad all-reduce)
"""
__version__ = "0.1.0"


def patch_reference_loss(loss_module):
    """Point the reference's ``loss`` module (torch/loss.py, imported by train.py as ``loss_util``) at the CUDA ops of
    this package: ``compute_normals_sparse`` (loss.py:285) and ``compute_2dcolor_loss`` (loss.py:246).  Call once after
    ``import loss as loss_util``; signatures and results are the reference's."""
    from . import losses, normals
    loss_module.compute_normals_sparse = normals.compute_normals_sparse
    loss_module.compute_2dcolor_loss = losses.color_l1_loss
    return loss_module
