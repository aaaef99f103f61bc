#!/usr/bin/env python
"""Install recipe for the reference's Python sources (baseline arm + drop-in proof; NOT product source).

The reference has no installable package (no setup.py / pyproject.toml at its root; `pip install /root/reference`
has nothing to build), so this recipe copies the files the baseline legs execute -- unmodified -- from
`/root/reference/torch` into the git-ignored `baseline/_ref/torch/`:

    model.py loss.py style.py data_util.py scene_dataloader.py train.py test_scene.py test_scene_as_chunks.py
    category.npz utils/raycast_rgbd/raycast_rgbd.py utils/depth_utils/depth_utils.py (+ the packages' __init__.py)

`baseline/_ref/` is git-ignored but NOT gpurun-ignored: it travels to the GPU box like `oracle/_ref/*.so`, where
`/root/reference` does not exist.  Only `bench.py --impl reference`, the CPU baseline leg and `tests/` load anything
from it (through `baseline/ref_loader.py`); the product never does.

    python baseline/install_ref.py [--force]
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPSG_REFERENCE_TORCH", "/root/reference/torch")
DST = os.path.join(HERE, "_ref", "torch")
FILES = (
    "model.py", "loss.py", "style.py", "data_util.py", "scene_dataloader.py", "train.py", "test_scene.py",
    "test_scene_as_chunks.py", "category.npz",
    "utils/raycast_rgbd/__init__.py", "utils/raycast_rgbd/raycast_rgbd.py",
    "utils/depth_utils/__init__.py", "utils/depth_utils/depth_utils.py",
)


def sources_present():
    return all(os.path.isfile(os.path.join(SRC, f)) for f in FILES)


def installed():
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


def install(force=False):
    """Copy the files (unmodified).  Returns the destination directory."""
    if not sources_present():
        if installed():
            return DST  # GPU box: use what travelled with the snapshot
        raise FileNotFoundError("reference sources not found under %s and no copy under %s" % (SRC, DST))
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not force and os.path.isfile(d) and filecmp.cmp(s, d, shallow=False):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    init = os.path.join(DST, "utils", "__init__.py")  # the reference's `utils` is a namespace directory
    if not os.path.isfile(init):
        open(init, "w").close()
    return DST


if __name__ == "__main__":
    print("installed:", install(force="--force" in sys.argv))
