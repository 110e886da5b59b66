"""Runs one forward/backward case a few times (for `ncu -k regex:...`).
usage: python scripts/run_case.py fwd|bwd|fwd8 mode C cl H W oH oW [N]      (fwd8 = uint8 input)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolate_antialiasing_b200 import capi
kind, mode, C, cl, H, W, oH, oW = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), *map(int, sys.argv[5:9])
N = int(sys.argv[9]) if len(sys.argv) > 9 else max(1, int(2.5e8 // (C * (H * W + oH * oW) * 4)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
if kind in ("fwd", "fwd8"):
    x = torch.rand((N, C, H, W), generator=g, device=dev) * 255
    if kind == "fwd8": x = x.byte()
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    out = capi.resize_forward(x, (oH, oW), mode)
    for _ in range(4): capi.resize_forward(x, (oH, oW), mode, out=out)
else:
    go = torch.rand((N, C, oH, oW), generator=g, device=dev)
    if cl: go = go.contiguous(memory_format=torch.channels_last)
    for _ in range(5): capi.resize_backward(go, (N, C, H, W), mode)
torch.cuda.synchronize()
print("ok", N)
