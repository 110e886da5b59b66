"""cfg5 (BASELINE.json configs[4]): the scale sweep at N GPUs, full table.

    python scripts/bench_sweep.py [--groups fwd,mixed,uint8,bwd] [--quick]          # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_sweep.py

Dimensions, timing rule and roofline definition: scripts/sweep_lib.py (shared with bench.py, whose default line carries
the same sweep as `sub.cfg5`).  Weak scaling: every rank runs the whole sweep on its own batch (no collective on the
data path); a point's time is the MAX over ranks.  Prints one line per point and ONE JSON line with the aggregates."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    import torch.distributed as dist
    import sweep_lib
    from interpolate_antialiasing_b200 import capi

    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default="fwd,mixed,uint8,bwd")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pts = []
    for p in sweep_lib.point_list(tuple(args.groups.split(",")), quick=args.quick):
        r = sweep_lib.run_point(p, torch, capi, dev, gen, peak, dist, world)
        pts.append(r)
        if rank == 0:
            print(f"{sweep_lib.label(r):44s} N={r['N']:4d} {r['ms']*1e3:8.1f} us  {r['frac']*100:5.1f}% of peak", flush=True)
    capi.check_device(local_rank)
    if rank == 0:
        s = sweep_lib.summarize(pts, peak)
        line = {"metric": "Mpix/s and % of HBM peak, AA bilinear/bicubic resize fwd+bwd at 1/2/4/8 B200", "unit": "Mpix/s",
                "value": world * s["aggregate"]["Mpix_s"], "n_gpus": world, "scaling": "weak", "data": "synthetic",
                "config": {"workload": "cfg5 scale sweep (scripts/sweep_lib.py)", "groups": args.groups},
                "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", **s["aggregate"]}, "by_group": s["by_group"],
                "below_0.70": s["below_0.70"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
