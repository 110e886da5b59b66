"""ctypes/numpy front-end of the CPU oracle (oracle/aa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg -- never by interpolate_antialiasing_b200.  See the header of aa_oracle.c for
the reference file:line each function follows and for the parity-pinning status.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libaa_oracle.so")

BOX, TRIANGLE, CUBIC = 0, 1, 2
FILTERS = {"nearest": BOX, "box": BOX, "bilinear": TRIANGLE, "linear": TRIANGLE,
           "bicubic": CUBIC, "cubic": CUBIC}

_lib = None


def build(force=False):
    src = os.path.join(HERE, "aa_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B" if force else "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        for sfx in ("f32", "f64"):
            getattr(L, f"aa_oracle_interp_size_{sfx}").argtypes = [i64, i64, i32, i32]
            getattr(L, f"aa_oracle_tables_{sfx}").argtypes = [i64, i64, i32, i32, vp, vp, vp]
            getattr(L, f"aa_oracle_forward_{sfx}").argtypes = [vp] + [i64] * 8 + [vp] + [i64] * 6 + [i32, i32]
            getattr(L, f"aa_oracle_backward_nonaa_{sfx}").argtypes = [vp, i64, i64, i64, vp, i64, i64, i32]
            getattr(L, f"aa_oracle_backward_adjoint_{sfx}").argtypes = [vp, i64, i64, i64, vp, i64, i64, i32, i32]
        L.aa_oracle_set_axis_scale.argtypes = [ctypes.c_double]
        L.aa_oracle_set_scale_factors.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib = L
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle computes in float32/float64 only (reference dispatch "
                    f"aa_interpolation_impl.h:608-614), got {dtype}")


def _f(filter):
    return FILTERS[filter] if isinstance(filter, str) else int(filter)


class _axis_scale:
    """Context: the caller-provided scale factor of one axis (compute_scales_value), 0/None = not given."""

    def __init__(self, s):
        self.s = float(s or 0.0)

    def __enter__(self):
        lib().aa_oracle_set_axis_scale(ctypes.c_double(self.s))

    def __exit__(self, *a):
        lib().aa_oracle_set_axis_scale(ctypes.c_double(0.0))


class _scale_factors:
    def __init__(self, sf):
        self.sf = (0.0, 0.0) if sf is None else (float(sf[0] or 0.0), float(sf[1] or 0.0))

    def __enter__(self):
        lib().aa_oracle_set_scale_factors(ctypes.c_double(self.sf[0]), ctypes.c_double(self.sf[1]))

    def __exit__(self, *a):
        lib().aa_oracle_set_scale_factors(ctypes.c_double(0.0), ctypes.c_double(0.0))


def interp_size(in_size, out_size, filter, align_corners=False, dtype=np.float32, scale=None):
    with _axis_scale(scale):
        return getattr(lib(), f"aa_oracle_interp_size_{_sfx(dtype)}")(in_size, out_size, _f(filter), int(align_corners))


def tables(in_size, out_size, filter, align_corners=False, dtype=np.float32, scale=None):
    """-> (xmin int64[out], xsize int64[out], weights dtype[out, K]); scale = the axis' scale factor, if given"""
    dtype = np.dtype(dtype)
    K = interp_size(in_size, out_size, filter, align_corners, dtype, scale)
    xmin = np.empty(out_size, np.int64)
    xsize = np.empty(out_size, np.int64)
    w = np.empty((out_size, K), dtype)
    with _axis_scale(scale):
        k2 = getattr(lib(), f"aa_oracle_tables_{_sfx(dtype)}")(
            in_size, out_size, _f(filter), int(align_corners), xmin.ctypes.data, xsize.ctypes.data, w.ctypes.data)
    assert k2 == K
    return xmin, xsize, w


def dense_matrix(in_size, out_size, filter, align_corners=False, dtype=np.float32):
    """[out, in] dense weight matrix equivalent to the tables."""
    xmin, xsize, w = tables(in_size, out_size, filter, align_corners, dtype)
    m = np.zeros((out_size, in_size), w.dtype)
    for o in range(out_size):
        m[o, xmin[o]:xmin[o] + xsize[o]] = w[o, :xsize[o]]
    return m


def forward(x, output_size, filter, align_corners=False, channels_last_out=None, scale_factors=None):
    """x: numpy [N,C,H,W] (any strides, float32/float64).  Output memory format follows the input
    (aa_interpolation_impl.h:752): channels_last strides in -> channels_last strides out."""
    x = np.asarray(x)
    assert x.ndim == 4
    N, C, H, W = x.shape
    oH, oW = int(output_size[0]), int(output_size[1])
    es = x.dtype.itemsize
    isn, isc, ish, isw = (s // es for s in x.strides)
    if channels_last_out is None:
        channels_last_out = C > 1 and isc == 1 and isw == C
    if channels_last_out:
        buf = np.empty((N, oH, oW, C), x.dtype)
        out = buf.transpose(0, 3, 1, 2)
    else:
        out = np.empty((N, C, oH, oW), x.dtype)
    osn, osc, osh, osw = (s // es for s in out.strides)
    with _scale_factors(scale_factors):
        rc = getattr(lib(), f"aa_oracle_forward_{_sfx(x.dtype)}")(
            x.ctypes.data, N, C, H, W, isn, isc, ish, isw, out.ctypes.data, oH, oW, osn, osc, osh, osw,
            _f(filter), int(align_corners))
    assert rc == 0
    return out


def backward_nonaa(grad_out, input_size, align_corners=False):
    """The reference's exported linear_backward (NON-antialiased; aa_interpolation_backward_impl.h:80-108)."""
    g = np.ascontiguousarray(grad_out)
    N, C, oH, oW = g.shape
    H, W = int(input_size[-2]), int(input_size[-1])
    gi = np.empty((N, C, H, W), g.dtype)
    rc = getattr(lib(), f"aa_oracle_backward_nonaa_{_sfx(g.dtype)}")(
        g.ctypes.data, N * C, oH, oW, gi.ctypes.data, H, W, int(align_corners))
    assert rc == 0
    return gi


def backward_adjoint(grad_out, input_size, filter, align_corners=False, scale_factors=None):
    """True adjoint of forward(): Wh^T g Ww from the bit-exact tables."""
    g = np.ascontiguousarray(grad_out)
    N, C, oH, oW = g.shape
    H, W = int(input_size[-2]), int(input_size[-1])
    gi = np.empty((N, C, H, W), g.dtype)
    with _scale_factors(scale_factors):
        rc = getattr(lib(), f"aa_oracle_backward_adjoint_{_sfx(g.dtype)}")(
            g.ctypes.data, N * C, oH, oW, gi.ctypes.data, H, W, _f(filter), int(align_corners))
    assert rc == 0
    return gi
