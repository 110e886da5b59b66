"""cfg2 at shrinking batch sizes (what each rank sees under strong scaling): time per image and fraction of peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolate_antialiasing_b200 import capi
dev = torch.device("cuda", 0)
peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
g = torch.Generator(device=dev).manual_seed(0)
xfull = (torch.rand((256, 3, 1080, 1920), generator=g, device=dev) * 255).contiguous(memory_format=torch.channels_last)
for n in (256, 128, 64, 32, 16, 8):
    x = xfull[:n]
    out = capi.resize_forward(x, (224, 224), "linear")
    for _ in range(3):
        capi.resize_forward(x, (224, 224), "linear", out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for i in range(20):
        capi.resize_forward(x, (224, 224), "linear", out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1]) / 20
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    b = n * 3 * (1080 * 1920 + 224 * 224) * 4
    print(f"N={n:4d}  {ms*1e3:8.1f} us  median {per[10]*1e3:8.1f}  per image {ms*1e3/n:6.2f} us  frac {b/ms/1e6/peak:.3f}  (flag AA_STREAM_GRID_MUL={os.environ.get('AA_STREAM_GRID_MUL','2')})", flush=True)
