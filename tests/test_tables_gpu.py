"""K1 parity: the sm_100a table kernel against the oracle, through the C ABI (aa_build_tables).
Bar: bit-exact xmin/xsize AND bit-exact weights (integer/index work; weights are achievable too)."""
import random

import numpy as np
import pytest
import torch

from oracle import aa_oracle as O

pytestmark = pytest.mark.gpu

MODES = ["linear", "cubic", "nearest"]


def _check(capi, a, b, mode, align, tdt):
    ndt = np.float32 if tdt == torch.float32 else np.float64
    xmin, xsize, w = capi.build_tables(a, b, mode, align, tdt)
    oxmin, oxsize, ow = O.tables(a, b, mode, align, ndt)
    assert w.shape == ow.shape, (a, b, mode, align, tdt)
    assert np.array_equal(xmin.cpu().numpy(), oxmin), (a, b, mode, align, tdt)
    assert np.array_equal(xsize.cpu().numpy(), oxsize), (a, b, mode, align, tdt)
    assert np.array_equal(w.cpu().numpy(), ow), (a, b, mode, align, tdt, np.abs(w.cpu().numpy() - ow).max())


def test_tables_named_configs(cuda):
    from interpolate_antialiasing_b200 import capi
    named = [(906, 320), (438, 196), (1920, 224), (1080, 224), (3840, 512), (2160, 512), (512, 128), (64, 10)]
    for a, b in named:
        for mode in MODES:
            for align in (False, True):
                for tdt in (torch.float32, torch.float64):
                    _check(capi, a, b, mode, align, tdt)


def test_tables_random_sweep(cuda):
    from interpolate_antialiasing_b200 import capi
    rnd = random.Random(5)
    cases = [(7, 7), (5, 1), (1, 5), (1, 1), (400, 3), (3, 400), (2, 1000), (1000, 2), (4096, 4095), (4095, 4096)]
    cases += [(rnd.randint(1, 600), rnd.randint(1, 600)) for _ in range(120)]
    cases += [(rnd.randint(1000, 5000), rnd.randint(1, 3000)) for _ in range(20)]
    for a, b in cases:
        mode = rnd.choice(MODES)
        for align in (False, True):
            _check(capi, a, b, mode, align, rnd.choice([torch.float32, torch.float64]))


def test_interp_size_host_matches_oracle():
    # host-only query, no device needed -- but kept here so it runs on the box next to the kernel
    from interpolate_antialiasing_b200 import capi
    rnd = random.Random(9)
    for _ in range(500):
        a, b = rnd.randint(1, 5000), rnd.randint(1, 5000)
        mode = rnd.choice(MODES)
        align = rnd.random() < 0.5
        assert capi.interp_size(a, b, mode, align, capi.F32) == O.interp_size(a, b, mode, align, np.float32)
        assert capi.interp_size(a, b, mode, align, capi.F64) == O.interp_size(a, b, mode, align, np.float64)


def test_device_tables_equal_host_tables_and_scale_factors(cuda):
    """K1 (device) against the host-side integer tables used for launch planning, and both against the oracle, with the
    reference's `scale_factors` knob (aa_interpolation_impl.h:735,740-742 -> area_pixel_compute_scale) exercised."""
    from interpolate_antialiasing_b200 import capi
    rnd = random.Random(21)
    cases = [(1080, 224), (3840, 512), (512, 128), (128, 512), (97, 97)] + [(rnd.randint(1, 3000), rnd.randint(1, 2000)) for _ in range(60)]
    for a, b in cases:
        mode = rnd.choice(MODES)
        tdt = rnd.choice([torch.float32, torch.float64])
        ndt = np.float32 if tdt == torch.float32 else np.float64
        for scale in (None, b / a * rnd.uniform(0.8, 1.25)):
            xmin, xsize, w = capi.build_tables(a, b, mode, False, tdt, scale=scale)
            hm, hs = capi.host_tables(a, b, mode, False, capi.F32 if tdt == torch.float32 else capi.F64, scale)
            oxmin, oxsize, ow = O.tables(a, b, mode, False, ndt, scale)
            assert np.array_equal(xmin.cpu().numpy(), hm) and np.array_equal(xsize.cpu().numpy(), hs), (a, b, mode, scale)
            assert np.array_equal(hm, oxmin) and np.array_equal(hs, oxsize), (a, b, mode, scale)
            assert np.array_equal(w.cpu().numpy(), ow), (a, b, mode, scale)
