"""Accuracy of the kernels' own fp64 exp (csrc/snp_math.cuh exp_tbl) against a correctly rounded reference."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_table_exp_is_within_2_ulp():
    from social_navigation_pyenvs_b200 import _lib as L
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-40, 5, 400000), rng.uniform(-1e-3, 1e-3, 50000),
                        np.array([0.0, -0.0, 1.0, -1.0, 700.0, -700.0, np.log(2) / 128, -np.log(2) / 128])])
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    L.check(L.lib().snp_debug_exp(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(yd.data_ptr()), x.size,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    y = yd.cpu().numpy()
    ref = np.exp(np.clip(x, -700, 700).astype(np.longdouble))   # x87 extended: good to ~1e-19 relative
    ulp = np.abs((y.astype(np.longdouble) - ref) / np.spacing(ref.astype(np.float64)).astype(np.longdouble))
    assert float(ulp.max()) < 2.0, float(ulp.max())
    assert float(np.mean(ulp)) < 0.5
    # below the guarded range the result flushes to (signed) zero for every purpose of the force laws; NaN propagates
    xs = torch.tensor([-1e3, -1e4, -1e6, float("nan")], dtype=torch.float64).cuda()
    ys = torch.empty_like(xs)
    L.check(L.lib().snp_debug_exp(ctypes.c_void_p(xs.data_ptr()), ctypes.c_void_p(ys.data_ptr()), 4,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    ys = ys.cpu().numpy()
    assert np.all(np.abs(ys[:3]) < 1e-250) and np.isnan(ys[3])


def _debug_math(kind, x, y=None, outs=1):
    from social_navigation_pyenvs_b200 import _lib as L
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    yd = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).cuda() if y is not None else None
    od = torch.empty(outs * x.size, dtype=torch.float64, device="cuda")
    L.check(L.lib().snp_debug_math(kind, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(yd.data_ptr() if yd is not None else 0),
                                   ctypes.c_void_p(od.data_ptr()), x.size, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return od.cpu().numpy().reshape(outs, x.size)


def _ulp(got, ref):
    ref = ref.astype(np.longdouble)
    return np.abs((got.astype(np.longdouble) - ref) / np.spacing(np.abs(ref).astype(np.float64) + 1e-300).astype(np.longdouble))


def test_scaled_exp2_is_within_2_ulp():
    """exp2_scaled(t) = 2^(t / 2048): what every A exp(x / B) of the force laws goes through (Params folds A, 1/B, 2048/ln2 into t)."""
    rng = np.random.RandomState(1)
    t = np.concatenate([rng.uniform(-2.0e6, 2.0e6, 200000), rng.uniform(-120000, 30000, 400000), rng.uniform(-3, 3, 50000),
                        np.array([0.0, -0.0, 0.5, -0.5, 1.0, 2048.0, -2048.0, 1023.5, 2.0e6, -2.0e6])])
    got = _debug_math(0, t)[0]
    ref = np.exp2(t.astype(np.longdouble) / 2048)
    ulp = _ulp(got, ref)
    assert float(ulp.max()) < 2.0, float(ulp.max())
    assert float(np.mean(ulp)) < 0.5
    # below the guarded range the value flushes to (a denormal next to) zero.  (A NaN exponent does not come back as NaN -- the
    # integer exponent arithmetic cannot keep it -- but every force is the product of this factor and the separation vector the
    # NaN came from, so a NaN state still yields a NaN force.)
    ys = _debug_math(0, np.array([-2.2e6, -1e7, -1e9, -2.0 ** 31 - 4716.0, -1e30, -np.inf]))[0]
    assert np.all(np.abs(ys) < 1e-290)
    # ... for EVERY exponent below the guard (a robot walking away from the crowd sweeps this range continuously; round 2 once
    # returned 1e308 at isolated points of it, where the significand polynomial crossed zero)
    sweep = np.concatenate([np.linspace(-3.0e6, -2.0e6, 2000001), -2.0 ** 31 - np.linspace(0, 20000, 400001)])
    ys = _debug_math(0, sweep)[0]
    assert np.all(np.isfinite(ys)) and float(np.abs(ys).max()) < 1e-290
    hi = _debug_math(0, np.linspace(2.0e6, 3.0e6, 100001))[0]
    assert np.all(np.isfinite(hi)) and np.all(hi > 1e290)


def test_atan2_sincos_rsqrt_clamp():
    rng = np.random.RandomState(2)
    n = 300000
    x = np.concatenate([rng.normal(0, 1, n), rng.normal(0, 1e-6, 1000), np.array([1.0, -1.0, 0.0, 0.0, 0.0, 1e-280, -3.0, 2.0])])
    y = np.concatenate([rng.normal(0, 1, n), rng.normal(0, 1e3, 1000), np.array([0.0, 0.0, 1.0, -1.0, 0.0, 1e-280, 1e-17, -2.0])])  # |(x, y)| >= 1e-280 or exactly 0
    got = _debug_math(1, x, y)[0]
    ref = np.arctan2(y.astype(np.longdouble), x.astype(np.longdouble))
    ok = (x != 0) | (y != 0)
    assert float(_ulp(got[ok], ref[ok]).max()) < 2.5
    assert got[~ok].tolist() == [0.0] * int((~ok).sum())
    a = np.concatenate([rng.uniform(-np.pi - 1, np.pi + 1, n), np.array([0.0, np.pi, -np.pi, np.pi / 2, -np.pi / 2, np.pi / 4, 1e-9])])
    sc = _debug_math(2, a, outs=2)
    al = a.astype(np.longdouble)
    # absolute error in units of the last place of 1 near the zeros of sin / cos (the reduction a - n pi/2 is good to 2^-60)
    assert float(np.max(np.abs(sc[0].astype(np.longdouble) - np.sin(al)))) < 2.5e-16
    assert float(np.max(np.abs(sc[1].astype(np.longdouble) - np.cos(al)))) < 2.5e-16
    v = np.exp(rng.uniform(-60, 60, n))
    r = _debug_math(3, v)[0]
    assert float(_ulp(r, 1 / np.sqrt(v.astype(np.longdouble))).max()) < 1.01
    t = np.concatenate([rng.uniform(-2, 3, 10000), np.array([0.0, -0.0, 1.0, 1.0 - 2**-53, 1.0 + 2**-52, -1e-300, 1e-300, 5e-324])])
    c = _debug_math(4, t)[0]
    assert np.array_equal(c, np.clip(t, 0.0, 1.0))
