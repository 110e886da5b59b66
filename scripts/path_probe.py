"""fp32 sweep points under each execution path (AUTO vs forced streaming kernel): which kernel should AUTO pick?
usage: python scripts/path_probe.py [C] [cl]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import sweep_lib  # noqa: E402
from interpolate_antialiasing_b200 import capi  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cl = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
for mode in ("linear", "cubic"):
    for s in (0.333, 0.5, 0.625, 0.75, 0.9, 1.0, 1.5, 2.0):
        p = dict(group="fwd", mode=mode, C=C, cl=cl, sh=s, sw=s, dtype="f32")
        row = [f"{mode:6s} s={s:5.3f}"]
        for name, fl in (("auto", capi.FLAG_AUTO), ("stream", capi.FLAG_FORCE_STREAM)):
            try:
                r = sweep_lib.run_point(p, torch, capi, dev, gen, peak, flags=fl, min_bytes=5e8)
                row.append(f"{name} {r['frac']:.3f}")
            except capi.AAError as e:
                row.append(f"{name} n/a")
        print("  ".join(row), flush=True)
