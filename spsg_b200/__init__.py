"""Importable alias of the product package.

The product lives in ``self-supervised-scene-generation-with-semantic-segmentation_b200/`` (the
directory name the project layout prescribes); hyphens make that name un-importable, so this stub
points ``spsg_b200``'s package path at it: ``import spsg_b200.raycast_rgbd`` loads
``self-supervised-scene-generation-with-semantic-segmentation_b200/raycast_rgbd.py``.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200")
__path__ = [PACKAGE_DIR]

with open(_os.path.join(PACKAGE_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
