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
