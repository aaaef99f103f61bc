#!/usr/bin/env python
"""Build recipe for the *reference* raycaster extension (test infrastructure, not product).

Compiles the two reference translation units where they lie,

    /root/reference/torch/utils/raycast_rgbd/raycast_rgbd_cuda.cpp
    /root/reference/torch/utils/raycast_rgbd/raycast_rgbd_cuda_kernel.cu

(unmodified, never copied into this repository) into ``oracle/_ref/spsg_ref_raycast_cuda.so``.
The pybind module name is the reference's own ``TORCH_EXTENSION_NAME`` macro, set here to
``spsg_ref_raycast_cuda`` so that it cannot shadow the product's drop-in ``raycast_rgbd_cuda``.

Flags mirror what ``torch.utils.cpp_extension.CUDAExtension`` would pass for
``TORCH_CUDA_ARCH_LIST=10.0`` (reference ``setup.py:1-14`` sets none of its own): ``-O3`` host,
``-gencode arch=compute_100,code=sm_100``, **no** fast-math, FMA contraction on.  ``-lineinfo`` is added
for ncu; it does not change code generation.

``oracle/_ref/`` is git-ignored but NOT gpurun-ignored: the built ``.so`` travels to the GPU box,
the sources do not (``/root/reference`` does not exist there).  Only ``tests/``, ``smoke()`` and
``bench.py --impl reference`` may load the result.
"""
import os
import shlex
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SPSG_REFERENCE_SRC", "/root/reference/torch/utils/raycast_rgbd")
OUT_DIR = os.path.join(HERE, "_ref")
BUILD_DIR = os.path.join(HERE, "_build")
MODULE = "spsg_ref_raycast_cuda"
OUT_SO = os.path.join(OUT_DIR, MODULE + ".so")
# second reference extension on the training step: utils/depth_utils, same recipe
DEPTH_SRC = os.environ.get("SPSG_REFERENCE_DEPTH_SRC", os.path.join(os.path.dirname(REF_SRC), "depth_utils"))
DEPTH_MODULE = "spsg_ref_depth_utils_cuda"
DEPTH_SO = os.path.join(OUT_DIR, DEPTH_MODULE + ".so")


def ref_sources_present():
    return all(os.path.isfile(os.path.join(REF_SRC, f))
               for f in ("raycast_rgbd_cuda.cpp", "raycast_rgbd_cuda_kernel.cu"))


def up_to_date():
    if not os.path.isfile(OUT_SO):
        return False
    if not ref_sources_present():
        return True  # nothing to rebuild from (GPU box): use the shipped binary
    t = os.path.getmtime(OUT_SO)
    srcs = [os.path.join(REF_SRC, f) for f in os.listdir(REF_SRC) if f.endswith((".cpp", ".cu", ".h"))]
    return all(os.path.getmtime(s) <= t for s in srcs) and os.path.getmtime(__file__) <= t


def run(cmd):
    print("+", " ".join(shlex.quote(c) for c in cmd), flush=True)
    subprocess.check_call(cmd)


def depth_sources_present():
    return all(os.path.isfile(os.path.join(DEPTH_SRC, f)) for f in ("depth_utils_cuda.cpp", "depth_utils_cuda_kernel.cu"))


def build_depth(force=False):
    """utils/depth_utils (bilateral filter, median fill, depth -> camera space, normals), unmodified, same flags."""
    if not depth_sources_present():
        if os.path.isfile(DEPTH_SO):
            return DEPTH_SO
        raise FileNotFoundError("reference sources not found under %s" % DEPTH_SRC)
    if not force and os.path.isfile(DEPTH_SO) and os.path.getmtime(DEPTH_SO) >= max(
            os.path.getmtime(os.path.join(DEPTH_SRC, f)) for f in os.listdir(DEPTH_SRC) if f.endswith((".cpp", ".cu", ".h"))):
        return DEPTH_SO
    return _compile(DEPTH_SRC, "depth_utils_cuda.cpp", "depth_utils_cuda_kernel.cu", DEPTH_MODULE, DEPTH_SO, "depth")


def build(force=False):
    if not force and up_to_date():
        return OUT_SO
    if not ref_sources_present():
        raise FileNotFoundError("reference sources not found under %s" % REF_SRC)
    return _compile(REF_SRC, "raycast_rgbd_cuda.cpp", "raycast_rgbd_cuda_kernel.cu", MODULE, OUT_SO, "ref")


def _compile(src_dir, cpp_name, cu_name, module, out_so, tag):
    import torch  # noqa: F401  (heavy import only when really building)
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(BUILD_DIR, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-isystem", p]
    defs = ["-DTORCH_EXTENSION_NAME=" + module, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    obj_cu = os.path.join(BUILD_DIR, tag + "_kernel.o")
    obj_cpp = os.path.join(BUILD_DIR, tag + "_shim.o")
    REF_SRC, OUT_SO = src_dir, out_so
    compat = ["-I", os.path.join(HERE, "compat")] if tag == "depth" else []  # see oracle/compat/torch/extension.h
    run([nvcc, "-c", os.path.join(REF_SRC, cu_name), "-o", obj_cu,
         *compat, "-I", REF_SRC, *inc, *defs,
         "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
         "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
         "--expt-relaxed-constexpr", "--compiler-options", "-fPIC", "-O3", "-lineinfo", "-w",
         "-gencode=arch=compute_100,code=sm_100", "-std=c++17"])
    run(["g++", "-c", os.path.join(REF_SRC, cpp_name), "-o", obj_cpp,
         "-I", REF_SRC, *inc, *defs, "-fPIC", "-O3", "-std=c++17", "-w"])
    libs = []
    for p in ce.library_paths("cuda"):
        libs += ["-L" + p, "-Wl,-rpath," + p]
    run(["g++", "-shared", obj_cpp, obj_cu, "-o", OUT_SO, *libs,
         "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"])
    return OUT_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_depth(force="--force" in sys.argv))
