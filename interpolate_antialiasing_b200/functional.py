"""autograd wiring, mirroring how the reference's harness pairs forward and backward
(/root/reference/test.py:123-157: get_proto_downsample_function / ProtoDownsample)."""
import torch

_MODES = {"bilinear": "linear", "linear": "linear", "bicubic": "cubic", "cubic": "cubic", "nearest": "nearest", "box": "nearest"}


class _AAResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, mode, align_corners):
        import interpolate_antialiasing_b200 as aa
        ctx.mode, ctx.align, ctx.size, ctx.ishape = mode, align_corners, tuple(size), tuple(x.shape)
        ctx.in_dtype = x.dtype
        return getattr(aa, mode + "_forward")(x, size, align_corners)

    @staticmethod
    def backward(ctx, grad_output):
        import interpolate_antialiasing_b200 as aa
        g = getattr(aa, ctx.mode + "_backward")(grad_output, ctx.size, ctx.ishape, ctx.align)
        if g.dtype != ctx.in_dtype and ctx.in_dtype.is_floating_point:
            g = g.to(ctx.in_dtype)
        return g, None, None, None


def aa_resize(x, size, mode="bilinear", align_corners=False):
    """Differentiable anti-aliased resize of a CUDA [N,C,H,W] tensor to `size` = (oH, oW)."""
    return _AAResizeFn.apply(x, tuple(size), _MODES[mode], bool(align_corners))


class AAResize(torch.nn.Module):
    """nn.Module wrapper (the reference's ProtoDownsample, test.py:149-157, but batched)."""

    def __init__(self, size, mode="bilinear", align_corners=False):
        super().__init__()
        self.size, self.mode, self.align_corners = tuple(size), mode, align_corners

    def forward(self, x):
        squeeze = x.dim() == 3
        out = aa_resize(x[None] if squeeze else x, self.size, self.mode, self.align_corners)
        return out[0] if squeeze else out


# ---- torch.library registration: the op as a first-class custom operator -------------------------
# `torch.ops.aa_b200.resize(x, [oH, oW], mode, align_corners)` carries a fake (meta) kernel and an
# autograd formula, so it traces under torch.compile / torch.export without graph breaks.  The real
# kernels are the same C-ABI calls as above; nothing here computes on the CPU.
_FWD = {"linear": "linear_forward", "cubic": "cubic_forward", "nearest": "nearest_forward"}
_BWD = {"linear": "linear_backward", "cubic": "cubic_backward", "nearest": "nearest_backward"}


def _out_dtype(x):
    return torch.float64 if x.dtype == torch.float64 else torch.float32


@torch.library.custom_op("aa_b200::resize", mutates_args=(), device_types="cuda")
def _resize_op(x: torch.Tensor, size: list[int], mode: str, align_corners: bool) -> torch.Tensor:
    import interpolate_antialiasing_b200 as aa
    return getattr(aa, _FWD[_MODES[mode]])(x, size, align_corners)


@_resize_op.register_fake
def _(x, size, mode, align_corners):
    fmt = torch.channels_last if (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()) \
        else torch.contiguous_format
    return torch.empty((x.shape[0], x.shape[1], size[0], size[1]), dtype=_out_dtype(x), device=x.device, memory_format=fmt)


@torch.library.custom_op("aa_b200::resize_backward", mutates_args=(), device_types="cuda")
def _resize_bwd_op(grad: torch.Tensor, size: list[int], input_size: list[int], mode: str, align_corners: bool) -> torch.Tensor:
    import interpolate_antialiasing_b200 as aa
    return getattr(aa, _BWD[_MODES[mode]])(grad, size, input_size, align_corners)


@_resize_bwd_op.register_fake
def _(grad, size, input_size, mode, align_corners):
    fmt = torch.channels_last if (grad.is_contiguous(memory_format=torch.channels_last) and not grad.is_contiguous()) \
        else torch.contiguous_format
    return torch.empty(tuple(input_size), dtype=grad.dtype, device=grad.device, memory_format=fmt)


def _setup_ctx(ctx, inputs, output):
    x, size, mode, align_corners = inputs
    ctx.size, ctx.mode, ctx.align, ctx.ishape, ctx.in_dtype = list(size), mode, align_corners, list(x.shape), x.dtype


def _bwd(ctx, grad):
    g = torch.ops.aa_b200.resize_backward(grad.contiguous() if not (grad.is_contiguous() or grad.is_contiguous(memory_format=torch.channels_last)) else grad,
                                          ctx.size, ctx.ishape, ctx.mode, ctx.align)
    if g.dtype != ctx.in_dtype and ctx.in_dtype.is_floating_point:
        g = g.to(ctx.in_dtype)
    return g, None, None, None


_resize_op.register_autograd(_bwd, setup_context=_setup_ctx)
