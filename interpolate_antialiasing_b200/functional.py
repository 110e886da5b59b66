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
