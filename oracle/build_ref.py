"""Compile the UNMODIFIED reference extension into oracle/_ref/ (test infrastructure only).

The reference path is one translation unit (step_two_dot_two/extension_interpolate.cpp, which
includes its two headers); it is compiled where it lies under /root/reference with
torch.utils.cpp_extension.load and the reference's own default flags (test.py:315-322: "-O3",
no -mfma, so nothing contracts).  Outputs go only into oracle/_ref/ (git-ignored, NOT
gpurun-ignored: the built .so travels to the GPU box, where /root/reference does not exist).

Nothing in the product path imports this; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs load the resulting module via oracle.ref_ext.load_ref().
"""
import os
import sys

REF_SRC = "/root/reference/step_two_dot_two/extension_interpolate.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "aa_ref_step_two_dot_two"


def build(verbose=False):
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(REF_SRC):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(REF_SRC):
        return so
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[REF_SRC], extra_cflags=["-O3"], build_directory=OUT,
         verbose=verbose)
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
