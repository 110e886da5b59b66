"""uint8 sweep points: tensor-core path vs the FP32-pipe kernels, with the wait-time counters (AA_VMMA_PROF=1)."""
import os
import sys
os.environ["AA_VMMA_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from interpolate_antialiasing_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
PEAK = 6531.6


def run(C, cl, s, mode, flags, tag):
    oh = ow = round(1024 * s)
    per = C * (1024 * 1024 + oh * ow * 4)
    N = int(1e9 // per) + 1
    x = torch.randint(0, 256, (N, C, 1024, 1024), generator=g, device=dev, dtype=torch.uint8)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    try:
        out = capi.resize_forward(x, (oh, ow), mode, False, flags)
    except capi.AAError as e:
        print(f"{tag:8s} {mode:6s} C={C} cl={cl} s={s}: {e}")
        return
    for _ in range(2):
        capi.resize_forward(x, (oh, ow), mode, False, flags, out=out)
    torch.cuda.synchronize()
    capi.debug_counters(0, reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    capi.resize_forward(x, (oh, ow), mode, False, flags, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c = capi.debug_counters(0)
    print(f"{tag:8s} {mode:6s} C={C} cl={cl} s={s} N={N}: {ms*1e3:8.1f} us frac {N*per/ms/1e6/PEAK:.3f}   "
          f"prodwait {c[2]/148:.0f} mma_wait_tile {c[5]/148:.0f} mma_wait_acc {c[4]/148:.0f} epi_wait {c[6]/148/16:.0f} "
          f"epi_tile {c[11]/148/16:.0f} hphase {c[9]/148/16:.0f} barrier {c[10]/148/16:.0f} life {c[14]/148/16:.0f}")


for mode in ("cubic", "linear"):
    for s in (0.125, 0.25, 0.333, 0.5, 0.75, 1.0):
        for C, cl in ((1, False), (3, True)):
            run(C, cl, s, mode, capi.FLAG_AUTO, "auto")
            run(C, cl, s, mode, capi.FLAG_VMMA, "vmma")
            run(C, cl, s, mode, capi.FLAG_FORCE_STREAM, "stream")
