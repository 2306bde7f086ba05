"""Dev tool: find the first sub-step at which an env of the bench crowd turns non-finite and print what happened around it."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
from social_navigation_pyenvs_b200 import CrowdEngine, _lib as L

np.set_printoptions(precision=17, linewidth=220)
inp = bench.build_inputs("4096x25_hsfm_ccso_walls_robot", 2000)
eng = CrowdEngine.from_reference_arrays(inp["model"], inp["states"], inp["goals"], walls=inp["walls"], safety=inp["safety"],
                                        consider_robot=True, all_params_equal=True, dtype=torch.float64)
E, N = inp["E"], inp["N"]
eng.action.copy_(torch.tensor(np.tile([[0.0], [1.0]], (1, E))))


def snapshot():
    return {k: getattr(eng, k).clone() for k in ("dyn", "goal_idx", "robot", "time_now")}


def restore(s):
    for k, v in s.items():
        getattr(eng, k).copy_(v)


def finite():
    return torch.isfinite(eng.dyn[:8]).all(0).view(E, N).all(1)


step = 0
while True:
    snap = snapshot()
    eng.step(None, bench.DT, n_substeps=20, pre_checks=True, post_checks=False, track_touch=True)
    ok = finite()
    if not bool(ok.all()):
        break
    step += 1
    if step > 400:
        print("no NaN in 400 steps"); sys.exit(0)
bad = (~ok).nonzero().flatten().cpu().numpy()
print(f"step {step}: {len(bad)} envs non-finite, first {bad[:10]}")
e = int(bad[0])
restore(snap)
for sub in range(20):
    before = eng.dyn.view(-1, E, N)[:, e].clone().cpu().numpy()
    rb = eng.robot[:, e].clone().cpu().numpy()
    eng.step(None, bench.DT, n_substeps=1, pre_checks=False, post_checks=False, track_touch=True)
    after = eng.dyn.view(-1, E, N)[:, e].clone().cpu().numpy()
    if not np.isfinite(after[:8]).all():
        j = np.where(~np.isfinite(after[:8]).all(0))[0]
        print(f"sub-step {sub}: env {e} humans {j} turn non-finite; robot px,py,vx,vy = {rb[:4]}")
        print("fields: px py vx vy th bvx bvy om dfx dfy")
        for jj in j[:3]:
            print(f"human {jj} before:", before[:, jj])
            print(f"human {jj} after :", after[:, jj])
        d = np.linalg.norm(before[:2].T[:, None] - before[:2].T[None], axis=-1)
        jj = int(j[0])
        print("distances from that human:", np.round(d[jj], 4))
        print("distance to robot:", np.linalg.norm(before[:2, jj] - rb[:2]))
        g = eng.goals.view(-1, 2, E, N)[:, :, e, jj].cpu().numpy()
        print("goals:", g, "goal_idx", int(eng.goal_idx.view(E, N)[e, jj]), "stat", eng.stat.view(-1, E, N)[:, e, jj].cpu().numpy())
        break

# replay that single sub-step on a one-env engine with pieces removed
st = eng.rows(inp["states"])  # NOTE: state AFTER the bad sub-step; rebuild the BEFORE state from `before`
rows = inp["states"][e:e + 1].copy()
rows[0, :N, 0] = before[0]; rows[0, :N, 1] = before[1]; rows[0, :N, 2] = before[4]
rows[0, :N, 3] = before[2]; rows[0, :N, 4] = before[3]; rows[0, :N, 5] = before[5]; rows[0, :N, 6] = before[6]; rows[0, :N, 7] = before[7]
rows[0, N, 0:2] = rb[0:2]; rows[0, N, 3:5] = rb[2:4]
gidx = eng.goal_idx.view(E, N)[e].cpu().numpy()
for name, walls, robot in (("all", inp["walls"], True), ("no walls", None, True), ("no robot", inp["walls"], False), ("neither", None, False)):
    r = rows if robot else rows[:, :N]
    s = inp["safety"][e:e + 1] if robot else inp["safety"][e:e + 1, :N]
    e1 = CrowdEngine.from_reference_arrays(inp["model"], r, inp["goals"][e:e + 1], walls=walls, safety=s, consider_robot=robot,
                                           all_params_equal=True, dtype=torch.float64, robot=None if robot else rows[:, N])
    e1.set_desired_force(np.stack([before[8], before[9]], -1)[None])
    e1.action.copy_(torch.tensor([[0.0], [1.0]]))
    e1.step(None, bench.DT, n_substeps=1, pre_checks=False, post_checks=False, track_touch=False)
    out = e1.dyn.view(-1, 1, N)[:, 0].cpu().numpy()
    print(name, "-> non-finite humans:", np.where(~np.isfinite(out[:8]).all(0))[0], "human", jj, "bv", out[5:7, jj])
