"""csrc/ref_trig.cuh (host build) against the libm the oracle and the compiled reference link: bit for bit."""
import numpy as np

import xgtest


def test_ref_trig_host_build_equals_libm(pkg):
    x = xgtest.trig_samples(400000)
    got = [np.empty_like(x) for _ in range(4)]
    pkg.lib().xgb_ref_trig_host(x.size, *[a.ctypes.data for a in [x] + got])
    want = xgtest.libm_trig(x)
    for name, a, b in zip(("sin", "cos", "sincos.sin", "sincos.cos"), got, want):
        bad = np.flatnonzero(a.view(np.uint64) != b.view(np.uint64))
        assert bad.size == 0, f"{name}: {bad.size} of {x.size} differ, first x={x[bad[0]]!r} {a[bad[0]]!r} vs {b[bad[0]]!r}"
    assert xgtest.libm_matches_ref_trig()


def test_ref_trig_site_host_build_equals_libm(pkg):
    """the clip kernel's single evaluation site (ref_trig_site / ref_sin_small) == libm sin / sincos, bit for bit"""
    x = xgtest.trig_samples(400000)
    got = [np.empty_like(x) for _ in range(4)]
    pkg.lib().xgb_ref_trig_site_host(x.size, *[a.ctypes.data for a in [x] + got])
    s, c, ss, sc = xgtest.libm_trig(x)
    for name, a, b in zip(("site.sin_only", "site.sincos.sin", "site.sincos.cos", "sin_small"), got, (s, ss, sc, s)):
        bad = np.flatnonzero(a.view(np.uint64) != b.view(np.uint64))
        assert bad.size == 0, f"{name}: {bad.size} of {x.size} differ, first x={x[bad[0]]!r} {a[bad[0]]!r} vs {b[bad[0]]!r}"
