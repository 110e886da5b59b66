"""cfg5 (BASELINE.json configs[4]): scale sweep 0.125x-2x (down + up), C in {1,3,4}, channels_first vs
channels_last, bilinear + bicubic, forward, at N GPUs (weak scaling: every rank runs the whole sweep on its
own batch shard; no collective on the data path).

    python scripts/bench_sweep.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_sweep.py

Per point: >= 1 GB of algorithmic traffic (input + output larger than L2), 2 warm-up + 5 timed calls, CUDA events
on the launching stream, median; across ranks the MAX of the medians.  Prints a table and ONE JSON line
(aggregate Mpix/s over all points and ranks, aggregate fraction of the measured HBM peak, per-point fractions).
This is a measurement helper next to bench.py (whose contract line is cfg2); same metric and unit."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from interpolate_antialiasing_b200 import capi

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
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    Hin = Win = 1024
    points = []
    tot_bytes = tot_pix = tot_ms = 0.0
    for mode in ("linear", "cubic"):
        for C in (1, 3, 4):
            for cl in (False, True):
                for s in (0.125, 0.25, 0.333, 0.5, 0.75, 1.0, 1.5, 2.0):
                    oh = ow = max(1, round(Hin * s))
                    per_img = C * (Hin * Win + oh * ow) * 4
                    N = int(1.0e9 // per_img) + 1
                    x = torch.rand((N, C, Hin, Win), generator=g, device=dev) * 255
                    if cl:
                        x = x.contiguous(memory_format=torch.channels_last)
                    out = capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO)
                    for _ in range(2):
                        capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO, out=out)
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
                    for a, b in ev:
                        a.record()
                        capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO, out=out)
                        b.record()
                    torch.cuda.synchronize()
                    ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
                    t = torch.tensor([ms], device=dev, dtype=torch.float64)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                    frac = N * per_img / (ms * 1e-3) / 1e9 / peak
                    points.append({"mode": mode, "C": C, "layout": "CL" if cl else "CF", "scale": s, "N": N, "ms": round(ms, 4),
                                   "frac": round(frac, 4)})
                    tot_bytes += N * per_img
                    tot_pix += N * (Hin * Win + oh * ow)
                    tot_ms += ms
                    if rank == 0:
                        print(f"{mode:6s} C={C} {'CL' if cl else 'CF'} s={s:5.3f} N={N:3d} {ms*1e3:8.1f} us  {frac*100:5.1f}% of peak", flush=True)
                    del x, out
    if rank == 0:
        fr = [p["frac"] for p in points]
        line = {"metric": "Mpix/s and % of HBM peak, AA bilinear/bicubic resize fwd+bwd at 1/2/4/8 B200", "unit": "Mpix/s",
                "value": world * tot_pix / 1e6 / (tot_ms * 1e-3), "n_gpus": world, "scaling": "weak", "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg5: scale sweep 0.125x-2x of [N,C,1024,1024], C in {1,3,4}, CF/CL, bilinear+bicubic, forward; "
                                       "N per point sized for >= 1 GB of traffic"},
                "roofline": {"bound": "hbm", "achieved": tot_bytes / (tot_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": tot_bytes / (tot_ms * 1e-3) / 1e9 / peak, "min_point_frac": min(fr), "max_point_frac": max(fr),
                             "points_at_or_above_0.70": sum(f >= 0.70 for f in fr), "points": len(fr)},
                "points": points}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
