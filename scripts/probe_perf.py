"""Quick device-time probe of the named configs (not the bench contract; used while tuning)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolate_antialiasing_b200 import capi

dev = torch.device("cuda", 0)
PEAK = 6531.6


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def run(name, x, osize, mode, flags, bytes_):
    try:
        out = capi.resize_forward(x, osize, mode, False, flags)
        med, best = timeit(lambda: capi.resize_forward(x, osize, mode, False, flags, out=out))
        print(f"{name:58s} med {med*1e3:9.1f} us  best {best*1e3:9.1f} us  {bytes_/med/1e6:8.1f} GB/s  {bytes_/med/1e6/PEAK*100:5.1f}% of peak", flush=True)
    except capi.AAError as e:
        print(f"{name:58s} ERROR {e}", flush=True)


which = sys.argv[1:] or ["cfg2", "cfg3", "cfg1", "cfg4", "torch"]
g = torch.Generator(device="cuda").manual_seed(0)
if "cfg2" in which:
    N = 64
    x = (torch.rand((N, 3, 1080, 1920), generator=g, device=dev) * 255).contiguous(memory_format=torch.channels_last)
    b = N * 3 * (1080 * 1920 + 224 * 224) * 4
    run("cfg2/4 fp32 CL 64x3x1080x1920->224 bilinear STREAM", x, (224, 224), "linear", capi.FLAG_FORCE_STREAM, b)
    run("cfg2/4 same GENERAL", x, (224, 224), "linear", capi.FLAG_FORCE_GENERAL, b)
    xcf = x.contiguous()
    run("cfg2/4 fp32 CF STREAM", xcf, (224, 224), "linear", capi.FLAG_FORCE_STREAM, b)
    if "torch" in which:
        med, best = timeit(lambda: torch.nn.functional.interpolate(x, size=(224, 224), mode="bilinear", antialias=True))
        print(f"{'torch F.interpolate(antialias=True) CL (incumbent)':58s} med {med*1e3:9.1f} us  {b/med/1e6:8.1f} GB/s")
        med, best = timeit(lambda: torch.nn.functional.interpolate(xcf, size=(224, 224), mode="bilinear", antialias=True))
        print(f"{'torch F.interpolate(antialias=True) CF (incumbent)':58s} med {med*1e3:9.1f} us  {b/med/1e6:8.1f} GB/s")
    del x, xcf
if "cfg3" in which:
    N = 32
    x = torch.randint(0, 256, (N, 3, 2160, 3840), generator=g, device=dev, dtype=torch.uint8)
    b = N * 3 * (2160 * 3840 * 1 + 512 * 512 * 4)
    run("cfg3/4 u8 CF 32x3x2160x3840->512 bicubic STREAM", x, (512, 512), "cubic", capi.FLAG_FORCE_STREAM, b)
    xf = x[:8].float()
    b2 = 8 * 3 * (2160 * 3840 * 4 + 512 * 512 * 4)
    run("cfg3 fp32 variant 8 imgs bicubic STREAM", xf, (512, 512), "cubic", capi.FLAG_FORCE_STREAM, b2)
    del x, xf
if "cfg1" in which:
    x = torch.rand((1, 3, 438, 906), generator=g, device=dev) * 255
    b = 3 * (438 * 906 + 196 * 320) * 4
    run("cfg1 fp32 CF 1x3x438x906->196x320 bilinear AUTO", x, (196, 320), "linear", capi.FLAG_AUTO, b)
    run("cfg1 same GENERAL", x, (196, 320), "linear", capi.FLAG_FORCE_GENERAL, b)
if "cfg4" in which:
    go = torch.rand((64, 3, 128, 128), generator=g, device=dev)
    b = 64 * 3 * (128 * 128 + 512 * 512) * 4
    gi = capi.resize_backward(go, (64, 3, 512, 512), "linear")
    med, best = timeit(lambda: capi.resize_backward(go, (64, 3, 512, 512), "linear"))
    print(f"{'cfg4 backward bilinear 64x3x128^2 -> 512^2':58s} med {med*1e3:9.1f} us  best {best*1e3:9.1f} us  {b/med/1e6:8.1f} GB/s  {b/med/1e6/PEAK*100:5.1f}% of peak")
    x = torch.rand((64, 3, 512, 512), generator=g, device=dev)
    run("cfg4 forward 64x3x512^2->128^2 bilinear AUTO", x, (128, 128), "linear", capi.FLAG_AUTO, b)

if "sweep" in which:
    # cfg5: scale sweep 0.125x-2x, C in {1,3,4}, channels_first vs channels_last, both filters; >= 1 GB per point
    Hin = Win = 1024
    rows = []
    for mode in ("linear", "cubic"):
        for C in (1, 3, 4):
            for cl in (False, True):
                for s in (0.125, 0.25, 0.333, 0.5, 0.75, 1.0, 1.5, 2.0):
                    oh = ow = max(1, round(Hin * s))
                    per_img = C * (Hin * Win + oh * ow) * 4
                    N = max(1, int(1.0e9 // per_img) + 1)
                    x = torch.rand((N, C, Hin, Win), generator=g, device=dev) * 255
                    if cl:
                        x = x.contiguous(memory_format=torch.channels_last)
                    try:
                        out = capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO)
                        med, best = timeit(lambda: capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO, out=out), iters=5, warm=2)
                        gbs = N * per_img / med / 1e6
                        print(f"sweep {mode:6s} C={C} {'CL' if cl else 'CF'} s={s:5.3f} N={N:3d} {med*1e3:9.1f} us {gbs:8.1f} GB/s {gbs/PEAK*100:5.1f}%", flush=True)
                    except capi.AAError as e:
                        print(f"sweep {mode} C={C} cl={cl} s={s} ERROR {e}", flush=True)
                    del x

if "tile" in which:
    # few-tap (tile kernel) cases only: forward 0.333x..2x bilinear/bicubic C=3 channels_first + backward of downsampling
    Hin = Win = 1024
    res = []
    for mode in ("linear", "cubic"):
        for s in (0.333, 0.5, 0.75, 1.0, 1.5, 2.0):
            oh = ow = round(Hin * s)
            per_img = 3 * (Hin * Win + oh * ow) * 4
            N = max(1, int(1.0e9 // per_img) + 1)
            x = torch.rand((N, 3, Hin, Win), generator=g, device=dev) * 255
            out = capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO)
            med, best = timeit(lambda: capi.resize_forward(x, (oh, ow), mode, False, capi.FLAG_AUTO, out=out), iters=5, warm=2)
            res.append(f"{mode[0]}{s:g}:{N * per_img / med / 1e6 / PEAK * 100:.1f}")
            del x, out
    for mode in ("linear", "cubic"):
        for (o, i) in ((128, 512), (512, 1024), (768, 1024)):
            per_img = 3 * (o * o + i * i) * 4
            N = max(1, int(3.0e8 // per_img) + 1)
            go = torch.rand((N, 3, o, o), generator=g, device=dev)
            capi.resize_backward(go, (N, 3, i, i), mode)
            med, best = timeit(lambda: capi.resize_backward(go, (N, 3, i, i), mode), iters=5, warm=2)
            res.append(f"b{mode[0]}{o}>{i}:{N * per_img / med / 1e6 / PEAK * 100:.1f}")
            del go
    print("tile " + os.environ.get("TAG", "") + " " + " ".join(res), flush=True)

if "tma" in which:
    N = 64
    x = (torch.rand((N, 3, 1080, 1920), generator=g, device=dev) * 255).contiguous(memory_format=torch.channels_last)
    b = N * 3 * (1080 * 1920 + 224 * 224) * 4
    ref = capi.resize_forward(x, (224, 224), "linear", False, capi.FLAG_FORCE_STREAM)
    run("cfg2/4 fp32 CL STREAM (LDG)", x, (224, 224), "linear", capi.FLAG_FORCE_STREAM, b)
    run("cfg2/4 fp32 CL STREAM (TMA)", x, (224, 224), "linear", capi.FLAG_FORCE_STREAM | capi.FLAG_STREAM_TMA, b)
    y = capi.resize_forward(x, (224, 224), "linear", False, capi.FLAG_FORCE_STREAM | capi.FLAG_STREAM_TMA)
    torch.cuda.synchronize()
    print("TMA vs LDG max abs diff", (y - ref).abs().max().item(), "equal", torch.equal(y, ref))
    xcf = x.contiguous()
    run("cfg2/4 fp32 CF STREAM (TMA)", xcf, (224, 224), "linear", capi.FLAG_FORCE_STREAM | capi.FLAG_STREAM_TMA, b)
    del x, xcf
    N = 32
    x = torch.randint(0, 256, (N, 3, 2160, 3840), generator=g, device=dev, dtype=torch.uint8)
    b = N * 3 * (2160 * 3840 * 1 + 512 * 512 * 4)
    ref = capi.resize_forward(x, (512, 512), "cubic", False, capi.FLAG_FORCE_STREAM)
    run("cfg3/4 u8 CF bicubic STREAM (LDG)", x, (512, 512), "cubic", capi.FLAG_FORCE_STREAM, b)
    run("cfg3/4 u8 CF bicubic STREAM (TMA)", x, (512, 512), "cubic", capi.FLAG_FORCE_STREAM | capi.FLAG_STREAM_TMA, b)
    y = capi.resize_forward(x, (512, 512), "cubic", False, capi.FLAG_FORCE_STREAM | capi.FLAG_STREAM_TMA)
    torch.cuda.synchronize()
    print("TMA vs LDG max abs diff (u8 cubic)", (y - ref).abs().max().item())
    del x


if "graph" in which:
    # launch-latency-scale configs: stream-launched back to back vs CUDA-graph replay (device time per call)
    def graph_time(fn, n=50):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(n):
                fn()
        gr.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    def burst_time(fn, n=200):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    x = torch.rand((1, 3, 438, 906), generator=g, device=dev) * 255
    out = capi.resize_forward(x, (196, 320), "linear")
    f1 = lambda: capi.resize_forward(x, (196, 320), "linear", False, capi.FLAG_AUTO, out=out)
    b1 = 3 * (438 * 906 + 196 * 320) * 4
    tg, tb = graph_time(f1), burst_time(f1)
    print(f"cfg1 forward: graph replay {tg*1e3:7.2f} us/call ({b1/tg/1e6:7.1f} GB/s), python back-to-back {tb*1e3:7.2f} us/call")
    go = torch.rand((64, 3, 128, 128), generator=g, device=dev)
    f4 = lambda: capi.resize_backward(go, (64, 3, 512, 512), "linear")
    b4 = 64 * 3 * (128 * 128 + 512 * 512) * 4
    tb = burst_time(f4)
    print(f"cfg4 backward: python back-to-back {tb*1e3:7.2f} us/call ({b4/tb/1e6:7.1f} GB/s, {b4/tb/1e6/PEAK*100:5.1f}% of peak)")

if "decode" in which:
    import interpolate_antialiasing_b200 as aa
    N = 128
    hwc = torch.randint(0, 256, (N, 1080, 1920, 3), generator=g, device=dev, dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    b = N * 3 * (1080 * 1920 * 1 + 224 * 224 * 2)
    med, best = timeit(lambda: aa.decode_resize_normalize(hwc, (224, 224), mean, std, "bilinear", torch.float16))
    print(f"{'decode-adjacent: HWC u8 128x1080x1920x3 -> CHW fp16 224^2 (1 kernel)':58s} med {med*1e3:9.1f} us  {b/med/1e6:8.1f} GB/s  {b/med/1e6/PEAK*100:5.1f}% of peak", flush=True)
    def unfused():
        x = hwc.permute(0, 3, 1, 2).float()
        y = torch.nn.functional.interpolate(x, size=(224, 224), mode="bilinear", antialias=True)
        m = torch.tensor(mean, device=dev).view(1, 3, 1, 1); s = torch.tensor(std, device=dev).view(1, 3, 1, 1)
        return ((y / 255.0 - m) / s).half().contiguous()
    med2, _ = timeit(unfused, iters=5, warm=2)
    print(f"{'same with torch ops (cast, F.interpolate AA, normalise, half)':58s} med {med2*1e3:9.1f} us  ({med2/med:.1f}x slower)", flush=True)

if "bwd" in which:
    for mode in ("linear", "cubic"):
        for (ih, oh) in ((512, 128), (1024, 512), (1024, 768), (512, 512), (512, 1024), (256, 1024)):
            C = 3
            per_img = C * (ih * ih + oh * oh) * 4
            N = max(1, int(4.0e8 // per_img))
            go = torch.rand((N, C, oh, oh), generator=g, device=dev)
            try:
                capi.resize_backward(go, (N, C, ih, ih), mode)
                med, best = timeit(lambda: capi.resize_backward(go, (N, C, ih, ih), mode), iters=5, warm=2)
                gbs = N * per_img / med / 1e6
                print(f"bwd {mode:6s} grad_out {oh}^2 -> grad_in {ih}^2 N={N:3d} {med*1e3:9.1f} us {gbs:8.1f} GB/s {gbs/PEAK*100:5.1f}%", flush=True)
            except capi.AAError as e:
                print("bwd ERROR", mode, ih, oh, e)
