import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from social_navigation_pyenvs_b200 import _lib as L
t = np.concatenate([np.linspace(-3.0e6, -2.0e6, 2000001), np.array([-2072734.5043227565])])
xd = torch.from_numpy(t).cuda(); od = torch.empty_like(xd)
L.check(L.lib().snp_debug_math(0, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(0), ctypes.c_void_p(od.data_ptr()), t.size, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
y = od.cpu().numpy()
print("non-finite:", int((~np.isfinite(y)).sum()), "max |y|:", np.nanmax(np.abs(y)), "at t =", t[np.nanargmax(np.abs(y))])
big = np.abs(y) > 1e-290
print("values above 1e-290:", int(big.sum()), t[big][:10], y[big][:10])
print("y at the failing t:", y[-1])
