"""Import the reference's unmodified Python modules from `baseline/_ref/torch` (baseline arm / tests only).

The reference's modules import, at module level, packages this image does not have (`imageio`, `plyfile`,
`skimage.color`) and its own CPU extensions (`utils.marching_cubes`, `utils.color_utils_cpu`) that none of the code
paths used here touch; empty stand-in modules are registered for exactly those names first.  The native raycaster
module name `raycast_rgbd_cuda` (reference `raycast_rgbd.py:7`) is bound to whichever implementation the caller asks
for: the product's drop-in (`spsg_b200.dropin.raycast_rgbd_cuda`) or the compiled reference extension
(`oracle/_ref/spsg_ref_raycast_cuda.so`).
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TORCH = os.path.join(HERE, "_ref", "torch")
_STUBS = ("imageio", "plyfile", "skimage", "skimage.color", "utils.marching_cubes", "utils.marching_cubes.marching_cubes",
          "utils.color_utils_cpu", "utils.color_utils_cpu.color_utils")


def available():
    return os.path.isfile(os.path.join(REF_TORCH, "model.py"))


def _stub_missing():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if "." not in name and importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError):
            pass
        m = types.ModuleType(name)
        m.__spsg_stub__ = True
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, m)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_module(name):
    """`model`, `loss`, `style`, `data_util` ... as the reference's scripts import them (top-level names)."""
    if not available():
        raise ImportError("baseline/_ref is not installed (python baseline/install_ref.py)")
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(REF_TORCH):
        return sys.modules[name]
    # the reference's `utils` directory must win over any other `utils` on sys.path while its modules load
    if "utils" not in sys.modules or not getattr(sys.modules["utils"], "__path__", [""])[0].startswith(REF_TORCH):
        pkg = types.ModuleType("utils")
        pkg.__path__ = [os.path.join(REF_TORCH, "utils")]
        sys.modules["utils"] = pkg
    _stub_missing()
    if name == "data_util" or name in ("loss", "style"):
        if "data_util" not in sys.modules or not getattr(sys.modules["data_util"], "__file__", "").startswith(REF_TORCH):
            _load("data_util", os.path.join(REF_TORCH, "data_util.py"))
        if name == "data_util":
            return sys.modules["data_util"]
    return _load(name, os.path.join(REF_TORCH, name + ".py"))


def bind_native(impl):
    """Make `import raycast_rgbd_cuda` (reference raycast_rgbd.py:7) resolve to `impl`: "ours" (the product's drop-in
    module over the C ABI) or "reference" (the compiled reference extension)."""
    if impl == "ours":
        from spsg_b200.dropin import raycast_rgbd_cuda as mod
    elif impl == "reference":
        from oracle import ref_driver
        mod = ref_driver.module()
    else:
        raise ValueError(impl)
    sys.modules["raycast_rgbd_cuda"] = mod
    return mod


def load_wrapper(impl):
    """The reference's unmodified `utils/raycast_rgbd/raycast_rgbd.py`, its native module bound to `impl`.  Returned
    under a per-implementation module name so that both bindings can coexist in one process."""
    if not available():
        raise ImportError("baseline/_ref is not installed (python baseline/install_ref.py)")
    bind_native(impl)
    return _load("spsg_ref_wrapper_" + impl, os.path.join(REF_TORCH, "utils", "raycast_rgbd", "raycast_rgbd.py"))
