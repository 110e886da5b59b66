"""autograd wiring, mirroring how the reference's harness pairs forward and backward
(/root/reference/test.py:123-157: get_proto_downsample_function / ProtoDownsample)."""
from typing import Optional

import torch

_MODES = {"bilinear": "linear", "linear": "linear", "bicubic": "cubic", "cubic": "cubic", "nearest": "nearest", "box": "nearest"}


def resolve_size(in_hw, size=None, scale_factor=None, recompute_scale_factor=None):
    """torch.nn.functional.interpolate's size/scale_factor rules -> (output_size, scale_factors or None).

    scale_factors is what reaches the table builder (the reference's `scale_factors`, aa_interpolation_impl.h:735,
    740-742): given when the caller passed scale_factor and did not ask to recompute it; the table scale is then
    1/scale_factor instead of in/out (area_pixel_compute_scale)."""
    if (size is None) == (scale_factor is None):
        raise ValueError("exactly one of size and scale_factor must be given")
    if size is not None:
        return (int(size[0]), int(size[1])), None
    sf = (float(scale_factor),) * 2 if not isinstance(scale_factor, (tuple, list)) else tuple(float(v) for v in scale_factor)
    if len(sf) != 2:
        raise ValueError("scale_factor must be a number or a pair")
    out = tuple(int(float(n) * f) for n, f in zip(in_hw, sf))  # floor(in * scale), compute_output_size
    return out, (None if recompute_scale_factor else sf)


class _AAResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, mode, align_corners, scale_factors):
        import interpolate_antialiasing_b200 as aa
        ctx.mode, ctx.align, ctx.size, ctx.ishape, ctx.sf = mode, align_corners, tuple(size), tuple(x.shape), scale_factors
        ctx.in_dtype = x.dtype
        return getattr(aa, mode + "_forward")(x, None if scale_factors else size, align_corners, scale_factors)

    @staticmethod
    def backward(ctx, grad_output):
        import interpolate_antialiasing_b200 as aa
        g = getattr(aa, ctx.mode + "_backward")(grad_output, None if ctx.sf else ctx.size, ctx.ishape, ctx.align, ctx.sf)
        if g.dtype != ctx.in_dtype and ctx.in_dtype.is_floating_point:
            g = g.to(ctx.in_dtype)
        return g, None, None, None, None


def aa_resize(x, size=None, mode="bilinear", align_corners=False, scale_factor=None, recompute_scale_factor=None):
    """Differentiable anti-aliased resize of a CUDA [N,C,H,W] tensor to `size` = (oH, oW), or by `scale_factor` with
    F.interpolate's semantics (recompute_scale_factor=True: only the output size is derived from it)."""
    osize, sf = resolve_size(x.shape[-2:], size, scale_factor, recompute_scale_factor)
    return _AAResizeFn.apply(x, osize, _MODES[mode], bool(align_corners), sf)


class AAResize(torch.nn.Module):
    """nn.Module wrapper (the reference's ProtoDownsample, test.py:149-157, but batched)."""

    def __init__(self, size=None, mode="bilinear", align_corners=False, scale_factor=None, recompute_scale_factor=None):
        super().__init__()
        self.size = None if size is None else tuple(size)
        self.mode, self.align_corners = mode, align_corners
        self.scale_factor, self.recompute_scale_factor = scale_factor, recompute_scale_factor

    def forward(self, x):
        squeeze = x.dim() == 3
        out = aa_resize(x[None] if squeeze else x, self.size, self.mode, self.align_corners, self.scale_factor,
                        self.recompute_scale_factor)
        return out[0] if squeeze else out


# ---- torch.library registration: the op as a first-class custom operator -------------------------
# `torch.ops.aa_b200.resize(x, [oH, oW], mode, align_corners)` carries a fake (meta) kernel and an
# autograd formula, so it traces under torch.compile / torch.export without graph breaks.  The real
# kernels are the same C-ABI calls as above; nothing here computes on the CPU.
_FWD = {"linear": "linear_forward", "cubic": "cubic_forward", "nearest": "nearest_forward"}
_BWD = {"linear": "linear_backward", "cubic": "cubic_backward", "nearest": "nearest_backward"}


def _out_dtype(x, mode="linear"):
    if x.dtype == torch.uint8 and _MODES[mode] == "nearest":
        return torch.uint8  # the box filter keeps uint8 (aa_interpolation_impl.h:566-570, :615-619)
    return torch.float64 if x.dtype == torch.float64 else torch.float32


@torch.library.custom_op("aa_b200::resize", mutates_args=(), device_types="cuda")
def _resize_op(x: torch.Tensor, size: list[int], mode: str, align_corners: bool,
               scale_factors: Optional[list[float]] = None) -> torch.Tensor:
    import interpolate_antialiasing_b200 as aa
    out = getattr(aa, _FWD[_MODES[mode]])(x, None if scale_factors else size, align_corners, scale_factors)
    assert list(out.shape[-2:]) == list(size), "size must be floor(in * scale_factor) when scale_factors is given"
    return out


@_resize_op.register_fake
def _(x, size, mode, align_corners, scale_factors=None):
    fmt = torch.channels_last if (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()) \
        else torch.contiguous_format
    return torch.empty((x.shape[0], x.shape[1], size[0], size[1]), dtype=_out_dtype(x, mode), device=x.device, memory_format=fmt)


@torch.library.custom_op("aa_b200::resize_backward", mutates_args=(), device_types="cuda")
def _resize_bwd_op(grad: torch.Tensor, size: list[int], input_size: list[int], mode: str, align_corners: bool,
                   scale_factors: Optional[list[float]] = None) -> torch.Tensor:
    import interpolate_antialiasing_b200 as aa
    return getattr(aa, _BWD[_MODES[mode]])(grad, None if scale_factors else size, input_size, align_corners, scale_factors)


@_resize_bwd_op.register_fake
def _(grad, size, input_size, mode, align_corners, scale_factors=None):
    fmt = torch.channels_last if (grad.is_contiguous(memory_format=torch.channels_last) and not grad.is_contiguous()) \
        else torch.contiguous_format
    return torch.empty(tuple(input_size), dtype=grad.dtype, device=grad.device, memory_format=fmt)


def _setup_ctx(ctx, inputs, output):
    x, size, mode, align_corners, scale_factors = inputs
    ctx.size, ctx.mode, ctx.align, ctx.ishape, ctx.in_dtype = list(size), mode, align_corners, list(x.shape), x.dtype
    ctx.sf = None if scale_factors is None else list(scale_factors)


def _bwd(ctx, grad):
    g = torch.ops.aa_b200.resize_backward(grad.contiguous() if not (grad.is_contiguous() or grad.is_contiguous(memory_format=torch.channels_last)) else grad,
                                          ctx.size, ctx.ishape, ctx.mode, ctx.align, ctx.sf)
    if g.dtype != ctx.in_dtype and ctx.in_dtype.is_floating_point:
        g = g.to(ctx.in_dtype)
    return g, None, None, None, None


_resize_op.register_autograd(_bwd, setup_context=_setup_ctx)
