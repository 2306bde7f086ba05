import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from social_navigation_pyenvs_b200 import scenarios, sensors
E, N = int(sys.argv[1]), 25
sc = scenarios.ccso_synthetic(E, N, seed0=2000)
humans = sc["states"][:, :, [0, 1, 8]]
walls = scenarios.pack_walls(scenarios.EXAMPLE_WALLS)
pose = np.concatenate([sc["robot"][:, 0:2], np.full((E, 1), np.pi / 2)], 1)
ranges, hits = sensors.scan_batch(humans, walls, pose, 2 * np.pi, 360, 10.0)
n = min(E, 256)
r_ref, h_ref = oracle.laser(humans[:n], walls, pose[:n], 2 * np.pi, 360, 10.0)
bad = np.argwhere(hits[:n] != h_ref)
print("hit mismatches", len(bad), "of", h_ref.size, "range maxdiff", np.abs(ranges[:n] - r_ref).max())
for e, k in bad[:10]:
    print(e, k, "gpu", hits[e, k], ranges[e, k], "ref", h_ref[e, k], r_ref[e, k])
