"""cfg5 (BASELINE.json configs[4]) measurement library, shared by bench.py (the `sub.cfg5` object of the driver-run
line) and scripts/bench_sweep.py (the full table).

Dimensions: scale 0.125x-2x down+up, C in {1,3,4}, channels_first vs channels_last, bilinear + bicubic,
input 1024x1024, N per point sized so that input+output traffic is >= 1 GB (larger than the 126 MB L2), plus
  mixed   different scales on the two axes (down x up, up x down, ...),
  uint8   uint8 input -> fp32 output (the fused cast; tensor-core vertical pass where eligible),
  bwd     the adjoint: grad_out [N,C,oh,ow] -> grad_in [N,C,1024,1024].
Per point: 2 warm-up + 5 timed calls, CUDA events on the launching stream, median; across ranks the MAX.
Roofline fraction of a point = algorithmic bytes (input read once + output written once) / time / measured HBM peak.
"""
SCALES = (0.125, 0.25, 0.333, 0.5, 0.75, 1.0, 1.5, 2.0)
MIXED = ((0.25, 0.75), (0.75, 0.25), (0.5, 2.0), (2.0, 0.5), (0.125, 1.0), (1.5, 0.333))
HIN = WIN = 1024


def point_list(groups=("fwd", "mixed", "uint8", "bwd"), quick=False):
    pts = []
    if "fwd" in groups:
        for mode in ("linear", "cubic"):
            for C in (1, 3, 4):
                for cl in (False, True):
                    for s in SCALES:
                        pts.append(dict(group="fwd", mode=mode, C=C, cl=cl, sh=s, sw=s, dtype="f32"))
    if "mixed" in groups:
        for mode in ("linear", "cubic"):
            for cl in (False, True):
                for sh, sw in MIXED:
                    pts.append(dict(group="mixed", mode=mode, C=3, cl=cl, sh=sh, sw=sw, dtype="f32"))
    if "uint8" in groups:
        for mode in ("linear", "cubic"):
            for C, cl in ((1, False), (3, False), (3, True), (4, True)):
                for s in (0.125, 0.25, 0.333, 0.5, 0.75, 1.0, 2.0):
                    pts.append(dict(group="uint8", mode=mode, C=C, cl=cl, sh=s, sw=s, dtype="u8"))
    if "bwd" in groups:
        for mode in ("linear", "cubic"):
            for cl in (False, True):
                for s in SCALES:
                    pts.append(dict(group="bwd", mode=mode, C=3, cl=cl, sh=s, sw=s, dtype="f32"))
    if quick:
        pts = pts[::7]
    return pts


def run_point(p, torch, capi, dev, gen, peak, dist=None, world=1, flags=0, min_bytes=1.0e9):
    oh, ow = max(1, round(HIN * p["sh"])), max(1, round(WIN * p["sw"]))
    C = p["C"]
    ies = 1 if p["dtype"] == "u8" else 4
    per_img = C * (HIN * WIN * ies + oh * ow * 4)
    N = int(min_bytes // per_img) + 1
    if p["group"] == "bwd":
        x = torch.rand((N, C, oh, ow), generator=gen, device=dev)
    elif p["dtype"] == "u8":
        x = torch.randint(0, 256, (N, C, HIN, WIN), generator=gen, device=dev, dtype=torch.uint8)
    else:
        x = torch.rand((N, C, HIN, WIN), generator=gen, device=dev) * 255
    if p["cl"]:
        x = x.contiguous(memory_format=torch.channels_last)
    if p["group"] == "bwd":
        out = capi.resize_backward(x, (N, C, HIN, WIN), p["mode"], False, flags=flags)
        call = lambda: capi.resize_backward(x, (N, C, HIN, WIN), p["mode"], False, flags=flags, out=out)
    else:
        out = capi.resize_forward(x, (oh, ow), p["mode"], False, flags)
        call = lambda: capi.resize_forward(x, (oh, ow), p["mode"], False, flags, out=out)
    for _ in range(2):
        call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record()
        call()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del x, out
    r = dict(p)
    r.update(N=N, ms=round(ms, 4), bytes=N * per_img, pix=N * (HIN * WIN + oh * ow), frac=round(N * per_img / (ms * 1e-3) / 1e9 / peak, 4))
    return r


def label(p):
    s = f"{p['sh']:g}" if p["sh"] == p["sw"] else f"{p['sh']:g}x{p['sw']:g}"
    return f"{p['group']:5s} {p['mode']:6s} C={p['C']} {'CL' if p['cl'] else 'CF'} {p['dtype']:3s} s={s}"


def summarize(points, peak):
    """aggregate (time-weighted) and per-group fractions, worst point, the points below the 0.70 target"""
    def agg(ps):
        tb = sum(p["bytes"] for p in ps)
        tm = sum(p["ms"] for p in ps)
        fr = [p["frac"] for p in ps]
        worst = min(ps, key=lambda p: p["frac"])
        return {"frac": round(tb / (tm * 1e-3) / 1e9 / peak, 4), "points": len(ps), "points_ge_0.70": sum(f >= 0.70 for f in fr),
                "min_point": {"frac": worst["frac"], "what": label(worst)}, "Mpix_s": round(sum(p["pix"] for p in ps) / 1e6 / (tm * 1e-3), 1)}
    out = {"aggregate": agg(points), "by_group": {}}
    for g in sorted(set(p["group"] for p in points)):
        out["by_group"][g] = agg([p for p in points if p["group"] == g])
    out["below_0.70"] = [{"what": label(p), "frac": p["frac"]} for p in sorted(points, key=lambda p: p["frac"]) if p["frac"] < 0.70]
    return out
