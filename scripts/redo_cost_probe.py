"""What the drain kernel behind the fast float kernels costs (csrc/aa_redo.cu): the same calls with the default flags
and with AA_FLAG_ASSUME_FINITE (no drain launch), same box, same buffers, interleaved.
    python scripts/redo_cost_probe.py            -> one line per case: us per call (default / assume-finite / difference)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from interpolate_antialiasing_b200 import capi


def timed(call, n, reps=7):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            call()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n)
    return sorted(ts)[len(ts) // 2]


def graph_us(call, n=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        return timed(g.replay, n)


def main():
    dev = torch.device("cuda", 0)
    cases = [
        ("cfg1 [1,3,438,906]->(196,320) linear fwd (stream)", (1, 3, 438, 906), (196, 320), "linear", False, "fwd", 200),
        ("cfg4 [64,3,128,128]->[64,3,512,512] linear bwd (tile)", (64, 3, 128, 128), (512, 512), "linear", False, "bwd", 100),
        ("band [16,3,1024,1024]->(768,768) linear fwd", (16, 3, 1024, 1024), (768, 768), "linear", False, "fwd", 50),
        ("tile [16,3,1024,1024]->(2048,2048) cubic fwd", (16, 3, 1024, 1024), (2048, 2048), "cubic", False, "fwd", 30),
        ("cfg2/8 [32,3,1080,1920] CL->(224,224) linear fwd (stream)", (32, 3, 1080, 1920), (224, 224), "linear", True, "fwd", 30),
        ("cfg2 [256,3,1080,1920] CL->(224,224) linear fwd (stream)", (256, 3, 1080, 1920), (224, 224), "linear", True, "fwd", 10),
    ]
    for name, shape, osize, mode, cl, what, n in cases:
        x = torch.rand(shape, device=dev) * 255
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        res = {}
        for tag, fl in (("default", capi.FLAG_AUTO), ("assume_finite", capi.FLAG_ASSUME_FINITE)):
            if what == "fwd":
                out = capi.resize_forward(x, osize, mode, False, fl)
                call = lambda: capi.resize_forward(x, osize, mode, False, fl, out=out)
            else:
                isz = (shape[0], shape[1]) + osize
                out = capi.resize_backward(x, isz, mode, False, flags=fl)
                call = lambda: capi.resize_backward(x, isz, mode, False, flags=fl, out=out)
            res[tag] = timed(call, n)
            if shape[0] == 1:
                res[tag + "_graph"] = graph_us(call)
        d = res["default"] - res["assume_finite"]
        line = f"{name:62s} default {res['default']:9.2f} us  assume_finite {res['assume_finite']:9.2f} us  drain costs {d:6.2f} us ({100 * d / res['assume_finite']:.1f} %)"
        if "default_graph" in res:
            line += f"   graph replay {res['default_graph']:.2f} / {res['assume_finite_graph']:.2f} us"
        print(line, flush=True)
        del x, out
    capi.check_device(0)


if __name__ == "__main__":
    main()
