"""cfg3 slice for ncu: uint8 CF 2160x3840 -> 512x512 bicubic, N images (default 8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolate_antialiasing_b200 import capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randint(0, 256, (N, 3, 2160, 3840), generator=g, device=dev, dtype=torch.uint8)
out = capi.resize_forward(x, (512, 512), "cubic")
for _ in range(4): capi.resize_forward(x, (512, 512), "cubic", out=out)
torch.cuda.synchronize()
print("ok")
