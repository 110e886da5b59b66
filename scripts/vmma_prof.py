"""Wait-time breakdown of the tensor-core kernel (AA_VMMA_PROF=1) on a cfg3 slice.  usage: python scripts/vmma_prof.py [N]"""
import os
import sys
os.environ["AA_VMMA_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from interpolate_antialiasing_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randint(0, 256, (N, 3, 2160, 3840), generator=g, device=dev, dtype=torch.uint8)
out = capi.resize_forward(x, (512, 512), "cubic")
for _ in range(3):
    capi.resize_forward(x, (512, 512), "cubic", out=out)
torch.cuda.synchronize()
capi.debug_counters(0, reset=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
capi.resize_forward(x, (512, 512), "cubic", out=out)
b.record()
torch.cuda.synchronize()
c = capi.debug_counters(0)
ms = a.elapsed_time(b)
items = N * 3 * 8 * 16
names = {1: "producer wait B slot", 2: "producer wait stage", 3: "MMA wait B", 4: "MMA wait acc release", 5: "MMA wait tile (TMA)",
         6: "epilogue wait acc ready (16 warps)", 7: "epilogue TMEM read (16 warps)", 11: "epilogue tile total (16 warps)", 8: "barrier before hphase (16 warps)", 9: "hphase (16 warps)", 10: "barrier after hphase (16 warps)",
         12: "producer lifetime", 13: "MMA lifetime", 14: "epilogue lifetimes (16 warps)"}
print(f"N={N} {ms*1e3:.1f} us, {items} items, {items/148:.1f} items/SM")
for k, n in names.items():
    div = 16 if "16 warps" in n else 1
    print(f"  [{k:2d}] {n:36s} {c[k]/148/div:12.0f} clk per SM-warp  = {c[k]/div/items:8.0f} clk/item")
