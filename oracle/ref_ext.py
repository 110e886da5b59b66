"""Loader for the UNMODIFIED reference extension compiled into oracle/_ref (see build_ref.py).

TEST INFRASTRUCTURE ONLY.  Returns the pybind module exporting the reference's own
linear_forward / cubic_forward / nearest_forward / linear_backward
(/root/reference/step_two_dot_two/extension_interpolate.cpp:46-51), or None when the prebuilt
.so is absent (e.g. a checkout where build() never ran).  /root/reference is NOT needed at load
time -- the GPU box only has the prebuilt file.
"""
import importlib.util
import os

from . import build_ref

_mod = None


def load_ref(build_if_missing=True):
    global _mod
    if _mod is not None:
        return _mod
    so = os.path.join(build_ref.OUT, build_ref.NAME + ".so")
    if not os.path.exists(so) and build_if_missing:
        build_ref.build()
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(build_ref.NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod
