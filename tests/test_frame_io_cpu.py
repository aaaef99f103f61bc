"""frame_io.load_frames (drop-in for data_util.load_frames, data_util.py:862-902) against the reference's own function on
small synthetic frame files.  The reference decodes with `imageio.imread`, which this image lacks; for these formats
imageio itself decodes with Pillow, so a three-line stand-in (`imread = numpy.array(PIL.Image.open(f))`) is registered
for it -- everything after decoding (resize, crop, scaling, intrinsic adjustment, batching) is the reference's code."""
import os
import sys
import types

import numpy as np
import pytest
import torch
from PIL import Image


def _write_scene(root, scene, frame_ids, size_wh, rng):
    for sub in ("depth", "color", "camera"):
        os.makedirs(os.path.join(root, scene, sub), exist_ok=True)
    w, h = size_wh
    for f in frame_ids:
        depth = rng.integers(0, 6000, size=(h, w)).astype(np.uint16)
        depth[rng.random((h, w)) < 0.1] = 0
        Image.fromarray(depth).save(os.path.join(root, scene, "depth", "%d.png" % f))
        color = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
        Image.fromarray(color).save(os.path.join(root, scene, "color", "%d.jpg" % f), quality=95)
        pose = np.eye(4) + rng.standard_normal((4, 4)) * 0.1
        intr = np.array([[1075.0 + f, 0, w / 2 - 0.5, 0], [0, 1076.0, h / 2 - 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        with open(os.path.join(root, scene, "camera", "%d.txt" % f), "w") as fh:
            for row in list(pose) + list(intr):
                fh.write(" ".join("%.6f" % v for v in row) + "\n")


def _reference_data_util():
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref not installed")
    shim = types.ModuleType("imageio")
    shim.imread = lambda f: np.array(Image.open(f))
    sys.modules["imageio"] = shim
    du = ref_loader.load_module("data_util")
    du.imageio = shim
    return du


@pytest.mark.parametrize("src_wh,dst_wh", [((64, 48), (40, 32)), ((80, 64), (80, 64)), ((96, 40), (32, 24))])
def test_load_frames_matches_reference(tmp_path, src_wh, dst_wh):
    from spsg_b200 import frame_io
    du = _reference_data_util()
    rng = np.random.default_rng(5)
    names = ["sceneA_room0__inc__3", "sceneB_room2__inc__1"]
    ids = {"sceneA": [3, 7, 9], "sceneB": [1, 2, 4]}
    for scene, fr in ids.items():
        _write_scene(str(tmp_path / "images"), scene, fr, src_wh, rng)
    os.makedirs(tmp_path / "frames")
    for name in names:
        with open(tmp_path / "frames" / (name.replace("__inc__", "__cmp__") + ".txt"), "w") as fh:
            fh.write("\n".join(str(i) for i in ids[name.split("_room")[0]]) + "\n")
    args = (names, None, str(tmp_path / "frames"), str(tmp_path / "images"), False, list(dst_wh), list(dst_wh), None, True, True)
    want = du.load_frames(*args, max_num_frames=2)
    for workers in (0, 4):
        got = frame_io.load_frames(*args, max_num_frames=2, num_workers=workers)
        for a, b in zip(got[:4], want[:4]):
            assert a.shape == b.shape and torch.equal(a, b)
        assert [list(f) for f in got[4]] == [list(f) for f in want[4]]
    # 'self' frame ids, depth only, too few frames
    got = frame_io.load_frames(names, None, "self", str(tmp_path / "images"), False, list(dst_wh), list(dst_wh), None, True, False)
    want = du.load_frames(names, None, "self", str(tmp_path / "images"), False, list(dst_wh), list(dst_wh), None, True, False)
    assert got[1] is None and want[1] is None and torch.equal(got[0], want[0]) and torch.equal(got[3], want[3])
    assert frame_io.load_frames(*args, max_num_frames=5) == (None, None, None, None, None)
