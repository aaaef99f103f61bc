"""The C-ABI library loads without a GPU and exports every symbol include/spsg_raycast.h declares; the drop-in
Python surface has the reference's names and signatures."""
import ctypes
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "spsg_raycast.h")).read()
    return sorted(set(re.findall(r"SPSG_API[^;(]*?\b(spsg_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spsg_b200 import _native
    syms = header_symbols()
    assert len(syms) >= 8
    assert sorted(_native.EXPORTS) == syms
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert "sm_100a" in _native.version()


def test_argument_validation_without_a_gpu():
    from spsg_b200 import _native as N
    p = N.make_params(320, 256, 5, 300, 45.45, 0.9, 64, 64, 128, 1, 1, 64, 0)
    assert N.workspace_bytes(p) > 0
    bad = N.make_params(0, 256, 5, 300, 45.45, 0.9, 64, 64, 128, 1, 1, 64, 0)
    assert N.lib.spsg_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"width" in N.lib.spsg_last_error()
    rc = N.lib.spsg_raycast_backward(ctypes.byref(p), *([None] * 12), 0, None)
    assert rc == 1 and b"NULL" in N.lib.spsg_last_error()


def test_dropin_surface_matches_reference_names():
    from spsg_b200 import raycast_rgbd, raycast_rgbd_cuda
    for name in ("forward", "backward", "construct_dense_sparse_mapping", "raycast_occ"):
        assert callable(getattr(raycast_rgbd_cuda, name))
    sig = inspect.signature(raycast_rgbd.RaycastRGBD.__init__)
    assert list(sig.parameters)[1:12] == ["batch_size", "dims3d", "width", "height", "depth_min", "depth_max",
                                          "thresh_sample_dist", "ray_increment", "max_num_frames",
                                          "max_num_locs_per_sample", "max_pixels_per_voxel"]
    assert sig.parameters["max_num_locs_per_sample"].default == 200000
    assert sig.parameters["max_pixels_per_voxel"].default == 64
    fsig = inspect.signature(raycast_rgbd.RaycastRGBD.forward)
    assert list(fsig.parameters)[1:] == ["locs", "vals_sdf", "vals_colors", "vals_normals", "vals_semantics",
                                         "view_matrix", "intrinsic_params"]
    osig = inspect.signature(raycast_rgbd.RaycastOcc.__init__)
    assert list(osig.parameters)[1:8] == ["batch_size", "dims3d", "width", "height", "depth_min", "depth_max",
                                          "ray_increment"]
    assert hasattr(raycast_rgbd, "RayCastRGBDFunction")


def test_reference_style_import_path():
    """train.py does `from utils.raycast_rgbd.raycast_rgbd import RaycastRGBD, RaycastOcc` and the wrapper does
    `import raycast_rgbd_cuda` (train.py:20-21, raycast_rgbd.py:7): both resolve with <package>/dropin on sys.path."""
    import subprocess
    import sys
    from spsg_b200 import PACKAGE_DIR
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from utils.raycast_rgbd.raycast_rgbd import RaycastRGBD, RaycastOcc; import raycast_rgbd_cuda;"
            "print(RaycastRGBD.__module__, raycast_rgbd_cuda.forward.__module__)"
            % (ROOT, os.path.join(PACKAGE_DIR, "dropin")))
    out = subprocess.check_output([sys.executable, "-c", code], text=True)
    assert "spsg_b200.raycast_rgbd" in out and "spsg_b200.raycast_rgbd_cuda" in out


def test_patch_reference_loss_module():
    """patch_reference_loss swaps the two loss.py functions on the hot path for the CUDA ops, keeping their signatures."""
    import types
    import spsg_b200
    from spsg_b200 import losses, normals
    fake = types.ModuleType("loss")
    fake.compute_normals_sparse = lambda sdf_locs, sdf_vals, dims, transform=None: None
    fake.compute_2dcolor_loss = lambda raycast_color, target_color, weight_color: None
    spsg_b200.patch_reference_loss(fake)
    assert fake.compute_normals_sparse is normals.compute_normals_sparse
    assert fake.compute_2dcolor_loss is losses.color_l1_loss
    assert list(inspect.signature(normals.compute_normals_sparse).parameters)[:4] == ["sdf_locs", "sdf_vals", "dims", "transform"]
    assert list(inspect.signature(losses.color_l1_loss).parameters) == ["raycast_color", "target_color", "weight_color"]


def test_flag_constants_match_the_header():
    """every SPSG_FLAG_* of include/spsg_raycast.h has the same value in the ctypes mirror, and no two flags share a bit"""
    from spsg_b200 import _native as N
    text = open(os.path.join(ROOT, "include", "spsg_raycast.h")).read()
    flags = {name: int(shift) for name, shift in re.findall(r"\b(SPSG_FLAG_\w+)\s*=\s*1u\s*<<\s*(\d+)", text)}
    assert len(flags) >= 8
    assert len(set(flags.values())) == len(flags)
    for name, shift in flags.items():
        assert getattr(N, name) == 1 << shift, name
