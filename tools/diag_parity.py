"""Dev diagnostic: per-column single-step error of the GPU step vs golden/oracle for named cases."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from oracle import OracleConfig
from helpers import *
from social_navigation_pyenvs_b200 import CrowdEngine, SFMS
np.set_printoptions(linewidth=200, precision=3)
names = sys.argv[2:] or traj_names()
dtype = torch.float32 if sys.argv[1] == "f32" else torch.float64
for name in names:
    d = load_traj(name); n = d["n"]
    ks = consecutive_pairs(d); ks = ks[:: max(1, len(ks) // 8)]
    for k in ks:
        S, G, D, rv = inputs_at(d, k)
        if dtype == torch.float32:
            S = S.astype(np.float32).astype(np.float64); D = D.astype(np.float32).astype(np.float64)
        if d["consider_robot"]:
            S[n, 0:2] = S[n, 0:2] + rv * float(d["dt"]); S[n, 3:5] = rv
        eng = CrowdEngine.from_reference_arrays(SFMS[int(d["type"])], S[None], G[None], walls=d["walls"], params=d["params"][None],
                                                safety=d["safety"][None, : S.shape[0]], consider_robot=d["consider_robot"],
                                                all_params_equal=d["all_equal"], dtype=dtype)
        eng.set_desired_force(D[None])
        eng.update_humans(0.0, float(d["dt"]))
        got = observed(eng.rows(S[None])[0], eng.desired_force()[0], n)
        cfg = OracleConfig(int(d["type"]), d["consider_robot"], d["all_equal"], False)
        S2, _, D2, F = oracle.update_humans(cfg, S[None], G[None], d["walls"], d["params"][None], d["safety"][None, : S.shape[0]], D[None], float(d["dt"]), 1, want_forces=True)
        ref = observed(S2[0], D2[0], n)
        e = rel_err(got, ref)
        if e[:, :10].max() > (1e-9 if dtype == torch.float64 else 1e-4):
            i = np.unravel_index(e[:, :10].argmax(), e[:, :10].shape)
            print(name, "step", int(d["steps"][k]), "max err", e[:, :10].max(), "at human/col", i, "got", got[i], "ref", ref[i])
            print("   col max:", e.max(0))
            print("   forces human", i[0], F[0, i[0]], "state", S[i[0]])
