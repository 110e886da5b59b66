"""Bring-up diagnostics for the tensor-core vertical pass (aa_vmma.cu): each case runs in its own process (a CUDA
fault poisons the context), structured inputs first, then random parity cases, against the oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "v_const": ("const1", (1, 1, 256, 512), (64, 512), "linear", False),
    "v_rand": ("rand", (1, 1, 256, 512), (64, 512), "linear", False),
    "v_cubic": ("rand", (1, 1, 256, 512), (61, 512), "cubic", False),
    "h_only": ("rand", (1, 1, 256, 512), (256, 128), "linear", False),
    "both": ("rand", (1, 1, 256, 512), (50, 70), "cubic", False),
    "n2": ("rand", (2, 1, 256, 512), (50, 70), "cubic", False),
    "c3": ("rand", (1, 3, 256, 512), (50, 70), "cubic", False),
    "n2c3": ("rand", (2, 3, 256, 512), (50, 70), "cubic", False),
    "big1": ("rand", (1, 1, 540, 960), (128, 128), "cubic", False),
    "big_n2c3": ("rand", (2, 3, 540, 960), (128, 128), "cubic", False),
    "big_cl": ("rand", (2, 3, 540, 960), (128, 128), "cubic", True),
    "cl_lin": ("rand", (2, 3, 540, 960), (100, 222), "linear", True),
    "wide": ("rand", (3, 1, 1000, 1504), (100, 1200), "cubic", False),
    "c4cl": ("rand", (2, 4, 300, 400), (64, 100), "linear", True),
}


def run_case(name):
    import torch
    from interpolate_antialiasing_b200 import capi
    from oracle import aa_oracle as O
    kind, shape, osize, mode, cl = CASES[name]
    rng = np.random.default_rng(0)
    if kind == "const1":
        xn = np.full(shape, 1, np.uint8)
    else:
        xn = rng.integers(0, 256, shape, dtype=np.uint8)
    x = torch.from_numpy(xn).cuda()
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    want = O.forward(x.float().cpu().numpy(), osize, mode)
    y = capi.resize_forward(x, osize, mode, False, capi.FLAG_VMMA)
    torch.cuda.synchronize()
    capi.check_device(0)
    got = y.cpu().numpy()
    err = np.abs(got.astype(np.float64) - want)
    ok = bool(np.all(err <= 1e-3 + 1e-5 * np.abs(want)))
    print(f"{name}: {shape} -> {osize} {mode} cl={cl}: max|err| {err.max():.3e} ok={ok}", flush=True)
    if not ok:
        bad = np.argwhere(err > 1e-3 + 1e-5 * np.abs(want))
        print(f"   {len(bad)} / {err.size} bad; first: {bad[:6].tolist()}")
        for b in bad[:6]:
            print(f"   at {tuple(b)}: got {got[tuple(b)]:.6f} want {want[tuple(b)]:.6f}")
        e2 = err.reshape(-1, err.shape[-2], err.shape[-1]).max(axis=0) > 1e-3
        print("   bad rows by oy%32:", np.bincount(np.nonzero(e2.any(axis=1))[0] % 32, minlength=32).tolist())
        print("   bad cols first 40:", np.nonzero(e2.any(axis=0))[0][:40].tolist())
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(run_case(sys.argv[1]))
    fails = 0
    for name in CASES:
        r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
        out = (r.stdout + r.stderr).strip().splitlines()
        if r.returncode != 0:
            fails += 1
            print(f"[{name}] exit {r.returncode}: " + " | ".join(out[-4:])[:600])
        else:
            print(out[-1] if out else f"[{name}] no output")
    print("ALL OK" if not fails else f"{fails} FAILURES")
    sys.exit(1 if fails else 0)
