#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the anti-aliased resize hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4]

One "step" = one pass of the hot path over one batch of synthetic images.  The headline workload
(BASELINE.json configs[1], "cfg2") is fp32 [256,3,1080,1920] channels_last -> (224,224) bilinear
antialias; each rank processes one such batch (weak scaling, no data-path collective; NCCL is used
only for the barrier and the max-over-ranks of the device time).

Prints ONE JSON line on rank 0:
  value     whole-job Mpix/s (pixels = N*(H*W + oH*oW), channels not counted) with inputs resident in HBM
  e2e       the same metric through the C ABI's host-buffer entry point (aa_resize_forward_host):
            pinned host input -> H2D -> kernel -> D2H of the result, all inside the timed region
  roofline  algorithmic bytes per launch / measured kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the UNMODIFIED reference extension (oracle/_ref) timed on this box's host cores on a
            bounded sample of the same workload (rank 0, N=1 only)

--impl reference times only the reference's CPU implementation (all host threads) and prints the
same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (N, C, H, W, oH, oW, mode, in_dtype, channels_last, kind)
    "cfg2": dict(N=256, C=3, H=1080, W=1920, oH=224, oW=224, mode="linear", dtype="f32", cl=True, kind="forward",
                 workload="cfg2: fp32 [256,3,1080,1920] channels_last -> (224,224) bilinear antialias, forward"),
    "cfg3": dict(N=128, C=3, H=2160, W=3840, oH=512, oW=512, mode="cubic", dtype="u8", cl=False, kind="forward",
                 workload="cfg3: uint8->fp32 [128,3,2160,3840] channels_first -> (512,512) bicubic antialias, forward"),
    "cfg4": dict(N=64, C=3, H=512, W=512, oH=128, oW=128, mode="linear", dtype="f32", cl=False, kind="backward",
                 workload="cfg4: backward of bilinear antialias [64,3,512,512] -> (128,128): grad_out [64,3,128,128] -> grad_in"),
}
METRIC = "Mpix/s (input+output pixels), AA bilinear/bicubic resize, and fraction of HBM peak"


def mpix(cfg, n_images=None):
    n = cfg["N"] if n_images is None else n_images
    return n * (cfg["H"] * cfg["W"] + cfg["oH"] * cfg["oW"]) / 1e6


def algorithmic_bytes(cfg, n_images=None):
    """input read once + output written once (no temp, tables or halo)."""
    n = cfg["N"] if n_images is None else n_images
    ies = 1 if cfg["dtype"] == "u8" else 4
    if cfg["kind"] == "forward":
        return n * cfg["C"] * (cfg["H"] * cfg["W"] * ies + cfg["oH"] * cfg["oW"] * 4)
    return n * cfg["C"] * (cfg["oH"] * cfg["oW"] + cfg["H"] * cfg["W"]) * 4


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(name):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process to the CPU cores local to its GPU (sysfs local_cpulist) so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's NUMA node.  Best effort; returns a note."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"cpus {txt} (GPU {bus} local)"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"
    return "unbound"


def make_inputs(cfg, dev, torch, n_images=None):
    n = cfg["N"] if n_images is None else n_images
    g = torch.Generator(device=dev).manual_seed(0)
    if cfg["kind"] == "backward":
        return torch.rand((n, cfg["C"], cfg["oH"], cfg["oW"]), generator=g, device=dev)
    shape = (n, cfg["C"], cfg["H"], cfg["W"])
    if cfg["dtype"] == "u8":
        x = torch.randint(0, 256, shape, generator=g, device=dev, dtype=torch.uint8)
    else:
        x = torch.rand(shape, generator=g, device=dev) * 255
    if cfg["cl"]:
        x = x.contiguous(memory_format=torch.channels_last)
    return x


def cpu_reference_rate(cfg, seconds, torch, threads=None):
    """Times the UNMODIFIED reference extension (oracle/_ref) on a bounded sample of cfg. -> dict"""
    from oracle.ref_ext import load_ref
    ref = load_ref(build_if_missing=False)
    if ref is None:
        return None
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = torch.get_num_threads()
    ns = 8 if cfg["H"] * cfg["W"] < 3e6 else 2
    g = torch.Generator().manual_seed(0)
    if cfg["kind"] == "backward":
        go = torch.rand((ns, cfg["C"], cfg["oH"], cfg["oW"]), generator=g)
        fn = lambda: ref.linear_backward(go, (cfg["oH"], cfg["oW"]), [ns, cfg["C"], cfg["H"], cfg["W"]], False)
        note = "reference linear_backward (non-AA arithmetic, SURVEY 0.2)"
    else:
        shape = (ns, cfg["C"], cfg["H"], cfg["W"])
        if cfg["dtype"] == "u8":
            xs = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
        else:
            xs = torch.rand(shape, generator=g) * 255
        if cfg["cl"]:
            xs = xs.contiguous(memory_format=torch.channels_last)
        f = ref.linear_forward if cfg["mode"] == "linear" else ref.cubic_forward
        # uint8: the reference's caller casts first (test.py:55,67); that cast is part of its path
        fn = (lambda: f(xs.float(), (cfg["oH"], cfg["oW"]), False)) if cfg["dtype"] == "u8" else (lambda: f(xs, (cfg["oH"], cfg["oW"]), False))
        note = "reference %s_forward%s" % (cfg["mode"], " incl. .float()" if cfg["dtype"] == "u8" else "")
    fn()
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter() - t0
    iters = max(2, min(200, int(seconds / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    dt = (time.perf_counter() - t0) / iters
    return {"value": mpix(cfg, ns) / dt, "unit": "Mpix/s", "cores": cores, "kind": "reference",
            "sample": f"{ns} of {cfg['N']} images of the same workload, {iters} passes, {note}, "
                      f"torch.get_num_threads()={cores}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')}, "
                      f"os.cpu_count()={os.cpu_count()}",
            "ms_per_pass": dt * 1e3, "images": ns}


def run_reference(args, cfg, rank, world):
    import torch
    if rank != 0:
        return
    line = {"metric": METRIC, "unit": "Mpix/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["workload"]}}
    from oracle.ref_ext import load_ref
    ref = load_ref(build_if_missing=False)
    if ref is None:
        line["unavailable"] = "oracle/_ref (compiled reference) is missing"
        print(json.dumps(line)); return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host thread it can
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    cores = torch.get_num_threads()
    ns = 8 if cfg["H"] * cfg["W"] < 3e6 else 2
    g = torch.Generator().manual_seed(0)
    if cfg["kind"] == "backward":
        go = torch.rand((ns, cfg["C"], cfg["oH"], cfg["oW"]), generator=g)
        fn = lambda: ref.linear_backward(go, (cfg["oH"], cfg["oW"]), [ns, cfg["C"], cfg["H"], cfg["W"]], False)
    else:
        shape = (ns, cfg["C"], cfg["H"], cfg["W"])
        xs = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8) if cfg["dtype"] == "u8" else torch.rand(shape, generator=g) * 255
        if cfg["cl"]:
            xs = xs.contiguous(memory_format=torch.channels_last)
        f = ref.linear_forward if cfg["mode"] == "linear" else ref.cubic_forward
        fn = (lambda: f(xs.float(), (cfg["oH"], cfg["oW"]), False)) if cfg["dtype"] == "u8" else (lambda: f(xs, (cfg["oH"], cfg["oW"]), False))
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    v = mpix(cfg, ns) / dt
    sample = (f"each step = {ns} of {cfg['N']} images of the workload through the unmodified reference extension "
              f"(oracle/_ref, -O3) on {cores} host threads; os.cpu_count()={os.cpu_count()}, "
              f"OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')}")
    line.update({"value": v, "ms_per_step": dt * 1e3,
                 "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--images", type=int, default=None, help="override the batch size (debug only; invalidates the headline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = dict(CONFIGS[args.config])
    if args.images:
        cfg["N"] = args.images
        cfg["workload"] += f" [DEBUG batch {args.images}]"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from interpolate_antialiasing_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x = make_inputs(cfg, dev, torch)
    osize = (cfg["oH"], cfg["oW"])
    if cfg["kind"] == "forward":
        out = capi.resize_forward(x, osize, cfg["mode"], False)
        step = lambda: capi.resize_forward(x, osize, cfg["mode"], False, capi.FLAG_AUTO, out=out)
    else:
        ishape = (cfg["N"], cfg["C"], cfg["H"], cfg["W"])
        step = lambda: capi.resize_backward(x, ishape, cfg["mode"], False)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    capi.launch_count(reset=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = capi.launch_count()
    total_ms = evs[0].elapsed_time(evs[-1])
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e and cfg["kind"] == "forward":
        numa_note = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else "single process, unbound"
        if cfg["cl"]:
            xh = torch.empty((cfg["N"], cfg["H"], cfg["W"], cfg["C"]), dtype=x.dtype, pin_memory=True).permute(0, 3, 1, 2)
            oh = torch.empty((cfg["N"], cfg["oH"], cfg["oW"], cfg["C"]), dtype=out.dtype, pin_memory=True).permute(0, 3, 1, 2)
        else:
            xh = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            oh = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        xh.copy_(x)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            capi.resize_forward_host(xh, oh, cfg["mode"], False, capi.FLAG_AUTO, device=local_rank)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            capi.resize_forward_host(xh, oh, cfg["mode"], False, capi.FLAG_AUTO, device=local_rank)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item()) / e2e_steps
        ok = bool(torch.allclose(oh.to(dev), out, rtol=1e-5, atol=1e-3))
        e2e = {"value": world * mpix(cfg) / dt, "unit": "Mpix/s", "h2d_bytes_per_step": xh.numel() * xh.element_size(),
               "d2h_bytes_per_step": oh.numel() * oh.element_size(), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "aa_resize_forward_host (C ABI, pinned host buffers)", "matches_device_result": ok,
               "host_affinity": numa_note}
        del xh, oh

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(cfg)
        kern_ms = total_ms / args.steps  # one launch per step; events on the launching stream
        achieved = ab / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * mpix(cfg) / (ms_per_step * 1e-3), "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "per_gpu_batch": cfg["N"], "sharding": "by image, no collective on the data path",
                       "l2": "inputs larger than L2 (%.2f GB read per step vs 126 MB L2)" % (ab / 1e9) if ab > 5e8 else
                             "working set fits L2: number is L2-warm, see DESIGN.md",
                       "path": "aa_resize_forward (C ABI) -> aa_stream_kernel" if cfg["kind"] == "forward" else "aa_resize_backward (C ABI)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(args.config), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ab, "kernel_ms": kern_ms,
                         "min_step_ms": per[0], "median_step_ms": per[len(per) // 2]},
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(cfg, args.cpu_seconds, torch)
            if cb:
                line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
