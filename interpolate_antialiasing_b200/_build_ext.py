"""In-tree build of the two native artefacts (both land in interpolate_antialiasing_b200/_build/):

  libaa_resize_b200.so   the C-ABI library: hand-written CUDA for sm_100a (nvcc, csrc/Makefile)
  aa_interp_b200.so      the torch C++ extension mirroring the reference's pybind module
                         (csrc/torch_binding.cpp), linked against the library above

nvcc cross-compiles without a GPU, so this runs on the CPU build container; the built files
travel to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(BUILD, "libaa_resize_b200.so")
EXT_NAME = "aa_interp_b200"
EXT = os.path.join(BUILD, EXT_NAME + ".so")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_library(verbose=False):
    subprocess.check_call(["make", "-C", CSRC, "-j4"] + ([] if verbose else ["-s"]))
    return LIB


def build_extension(verbose=False):
    src = os.path.join(CSRC, "torch_binding.cpp")
    hdr = os.path.join(ROOT, "include", "aa_resize.h")
    if _newer(EXT, [src, hdr]):
        return EXT
    import ctypes
    from torch.utils.cpp_extension import load
    os.makedirs(BUILD, exist_ok=True)
    ctypes.CDLL(LIB, mode=ctypes.RTLD_GLOBAL)  # so the freshly linked extension can be dlopen()ed by load()
    load(name=EXT_NAME, sources=[src], extra_include_paths=[os.path.join(ROOT, "include")],
         extra_cflags=["-O2"], extra_ldflags=[f"-L{BUILD}", "-laa_resize_b200"],
         with_cuda=True, build_directory=BUILD, verbose=verbose, is_python_module=False)
    return EXT


def build_all(verbose=False):
    build_library(verbose)
    build_extension(verbose)
    return LIB, EXT


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
