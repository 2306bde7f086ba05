"""Drop-in for the reference operator `update_humans_parallel` (social_gym/src/forces_parallel.py:184-284).

Same name, same argument order and meaning, same return value and the same in-place side effects on `goals` and
`agents_state` -- but it runs on the B200 through libsnp_b200.so, and it also accepts a leading env axis
(agents_state [E,N(+1),13], goals [E,N,G,2], agents_params [E,N,20], safety_space [E,N(+1)]) to step E independent
environments in one call.

Two keyword extensions select what the reference cannot express in this signature:
  semantics="serial" (default): the serial Python/NumPy path the operator shadows (motion_model_manager.py:369-373 +
        forces.py) -- the parity oracle named by the project;   semantics="numba": forces_parallel.py's own quirks
        ('<=' goal test fp:226, zeroed desired force fp:34-40, Guo wall force / n_walls fp:161, first-wins closest segment fp:252);
  desired_force: [.., N, 2] array carried between calls (serial path keeps a stale desired force inside the goal radius,
        forces.py:12-15); updated in place when given.
"""
import ctypes

import numpy as np

from . import _lib as L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def update_humans_parallel(type, agents_state, goals, obstacles, agents_params, dt, safety_space, all_params_equal=False,
                           last_is_robot=False, *, semantics="serial", desired_force=None, dtype="float64", n_substeps=1):
    if type < 0 or type > 8:
        raise ValueError(f"Type {type} does not exist for this implementation")  # forces_parallel.py:211
    if semantics not in ("serial", "numba"):
        raise ValueError("semantics must be 'serial' or 'numba'")
    batched = agents_state.ndim == 3
    st = agents_state if batched else agents_state[None]
    gl = goals if batched else goals[None]
    for name, arr in (("agents_state", st), ("goals", gl)):
        if arr.dtype != np.float64 or not arr.flags.c_contiguous:
            raise ValueError(f"{name} must be a C-contiguous float64 array (it is updated in place)")
    E, rows, width = st.shape
    if width != 13:
        raise ValueError("agents_state rows must have 13 columns [px,py,theta,vx,vy,bvx,bvy,omega,r,m,gx,gy,vd]")
    N, G = gl.shape[1], gl.shape[2]
    if rows != N + int(bool(last_is_robot)):
        raise ValueError(f"agents_state has {rows} rows, goals describe {N} humans, last_is_robot={last_is_robot}")
    params = np.ascontiguousarray(np.broadcast_to(np.asarray(agents_params, np.float64), (E, N, 20)))
    safety = np.ascontiguousarray(np.broadcast_to(np.asarray(safety_space, np.float64), (E, rows)))
    W = S = 0
    obs = None
    if obstacles is not None and np.size(obstacles) > 0:
        obs = np.ascontiguousarray(obstacles, np.float64)
        W, S = obs.shape[0], obs.shape[1]
    df = None
    if desired_force is not None:
        df = desired_force if batched else desired_force[None]
        if df.dtype != np.float64 or not df.flags.c_contiguous or df.shape != (E, N, 2):
            raise ValueError("desired_force must be a C-contiguous float64 array of shape [.., N, 2]")
    out = np.empty_like(st)
    rc = L.lib().snp_update_humans_parallel_host(int(type), E, N, G, _p(st), _p(gl), _p(obs), W, S, _p(params), float(dt), _p(safety),
                                                 int(bool(all_params_equal)), int(bool(last_is_robot)), int(semantics == "numba"),
                                                 L.SNP_F64 if dtype in ("float64", np.float64) else L.SNP_F32, int(n_substeps), _p(df), _p(out))
    L.check(rc)
    return out if batched else out[0]
