"""LaserSensor: host-side mirror of social_gym/src/sensors.py backed by the ray-cast kernel (csrc/snp_laser.cu).

Same constructor, `update_pose` and `get_laser_measurements` as the reference class for ONE sensor
(sensors.py:10-22,53-69), plus `scan_batch` for E sensors at once (arrays in, arrays out) and device-resident scans through
`CrowdEngine`-owned tensors (`scan_engine`).
"""
import ctypes
import math

import numpy as np

from . import _lib as L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _walls_array(walls):
    """Accepts the reference's wall objects (anything with a `.segments` dict, obstacle.py:26-32), a list of vertex lists,
    or an already packed [W,S,2,2] NaN padded array."""
    if walls is None:
        return None
    if isinstance(walls, np.ndarray):
        return walls if walls.size and walls.shape[0] else None
    walls = list(walls)
    if not walls:
        return None
    seg_lists = []
    for w in walls:
        if hasattr(w, "segments"):
            seg_lists.append([[list(s[0]), list(s[1])] for s in w.segments.values()])
        else:
            verts = [list(v) for v in w]
            seg_lists.append([[min(verts[i], verts[(i + 1) % len(verts)]), max(verts[i], verts[(i + 1) % len(verts)])]
                              for i in range(len(verts))])
    smax = max(len(s) for s in seg_lists)
    out = np.full((len(seg_lists), smax, 2, 2), np.nan)
    for i, segs in enumerate(seg_lists):
        out[i, : len(segs)] = np.asarray(segs, np.float64)
    return out


def scan_batch(humans, walls, pose, range_, samples, max_distance, robot_radius=0.0, dtype="float64", want_hits=True):
    """E sensors in one launch.  humans [E,N,3] = x,y,r; walls [W,S,2,2] NaN padded or None; pose [E,3] = x,y,yaw.
    Returns (ranges [E,samples] float64, hits [E,samples] int32 or None)."""
    if max_distance > 10:
        raise ValueError("Maxium distance for laser is 10 meters")  # sensors.py:13
    humans = np.ascontiguousarray(humans, np.float64)
    pose = np.ascontiguousarray(pose, np.float64)
    if np.any(pose[:, 2] > math.pi) or np.any(pose[:, 2] < -math.pi):
        raise ValueError("Angle passed ust be wrapped between [-pi,pi]")  # sensors.py:20
    E, N = humans.shape[0], humans.shape[1]
    w = _walls_array(walls)
    W, S = (0, 0) if w is None else (w.shape[0], w.shape[1])
    w = None if w is None else np.ascontiguousarray(w, np.float64)
    ranges = np.empty((E, samples))
    hits = np.empty((E, samples), np.int32) if want_hits else None
    rc = L.lib().snp_laser_host(E, N, _p(humans), _p(w), W, S, _p(pose), float(range_), int(samples), float(max_distance),
                                float(robot_radius), L.SNP_F64 if dtype in ("float64", np.float64) else L.SNP_F32, _p(ranges), _p(hits))
    L.check(rc)
    return ranges, hits


class LaserSensor:
    """social_gym/src/sensors.py:6-74 for one sensor.  The Gaussian noise of add_uncertainty (sensors.py:71-74) is applied on the
    host from the caller's global np.random stream, one draw per ray in ray order -- exactly the reference's draws.  (Batched scans
    take their noise on the device: EngineScanner(uncertainty=...).)"""

    def __init__(self, init_pos, init_yaw, range, samples, max_distance, uncertainty=None):
        self.range = range
        self.samples = samples
        if max_distance > 10:
            raise ValueError("Maxium distance for laser is 10 meters")
        self.max_distance = max_distance
        self.uncertainty = uncertainty
        self.update_pose(init_pos, init_yaw)

    def update_pose(self, position, yaw):
        if yaw > math.pi or yaw < -math.pi:
            raise ValueError("Angle passed ust be wrapped between [-pi,pi]")
        self.position = position
        self.yaw = yaw

    def get_laser_measurements(self, humans, walls):
        """humans: objects with .position/.radius (HumanAgent) or an [N,3] array; walls: reference wall objects, vertex lists
        or a packed array.  Returns {angle: range} like the reference."""
        if isinstance(humans, np.ndarray):
            h = np.asarray(humans, np.float64).reshape(-1, 3)
        else:
            h = np.array([[p.position[0], p.position[1], p.radius] for p in humans], np.float64).reshape(-1, 3)
        pose = np.array([[self.position[0], self.position[1], self.yaw]], np.float64)
        ranges, _ = scan_batch(h[None], walls, pose, self.range, self.samples, self.max_distance, want_hits=False)
        angles = np.linspace(self.yaw - (self.range / 2), self.yaw + (self.range / 2), self.samples)
        meas = ranges[0]
        if self.uncertainty:
            meas = np.array([max(min(np.random.normal(m, self.uncertainty), self.max_distance), 0) for m in meas])
        return dict(zip(angles, meas))


class EngineScanner:
    """Device-resident scans over a CrowdEngine's humans and walls with preallocated outputs and a cached argument block, so a
    scan costs one C call: `ranges, hits = scanner.scan(pose)` with pose a [3,E] device tensor (x, y, yaw) of the engine's dtype.
    The returned tensors are reused by the next scan."""

    def __init__(self, engine, range_, samples, max_distance, robot_radius=0.0, want_hits=True, uncertainty=None, seed=0):
        """uncertainty: LaserSensor's Gaussian range noise (sensors.py:71-74) applied on the device: every range becomes
        clip(N(range, uncertainty), 0, max_distance); ray k of env e draws from the Philox4x32-10 stream keyed by `seed` at
        counter (k, e, scan number), so scans are reproducible and successive scans independent.  (The reference draws from the
        caller's global np.random stream: same distribution, different numbers.)"""
        import torch
        if max_distance > 10:
            raise ValueError("Maxium distance for laser is 10 meters")
        self.engine, self.torch = engine, torch
        self.n_scans = 0
        self._pose_dev = None
        self.ranges = torch.empty((engine.E, samples), dtype=engine.dtype, device=engine.device)
        self.hits = torch.empty((engine.E, samples), dtype=torch.int32, device=engine.device) if want_hits else None
        a = L.SnpLaserArgs()
        a.E, a.N, a.samples = engine.E, engine.N, int(samples)
        a.dtype = L.SNP_F64 if engine.dtype == torch.float64 else L.SNP_F32
        a.px, a.py = engine.dyn[L.DYN_PX].data_ptr(), engine.dyn[L.DYN_PY].data_ptr()
        a.radius = engine.stat[L.STAT_R].data_ptr()
        a.walls = None if engine.walls is None else engine.walls.data_ptr()
        a.W, a.S, a.walls_per_env = engine.W, engine.S, engine.walls_per_env
        a.range, a.max_distance, a.robot_radius = float(range_), float(max_distance), float(robot_radius)
        a.ranges = self.ranges.data_ptr()
        a.hits = None if self.hits is None else self.hits.data_ptr()
        a.uncertainty = float(uncertainty) if uncertainty else 0.0
        a.noise_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.args, self.fn = a, L.lib().snp_laser

    def _launch(self, pose, ranges_ptr, hits_ptr):
        a = self.args
        a.pose, a.ranges, a.hits = pose.data_ptr(), ranges_ptr, hits_ptr
        a.noise_scan = self.n_scans
        self.n_scans += 1
        L.check(self.fn(ctypes.byref(a), ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)))

    def scan(self, pose):
        assert pose.is_contiguous() and pose.dtype == self.engine.dtype and pose.shape == (3, self.engine.E)
        self._launch(pose, self.ranges.data_ptr(), None if self.hits is None else self.hits.data_ptr())
        return self.ranges, self.hits

    def scan_host(self, pose_host, ranges_host, hits_host=None):
        """The same scan with HOST buffers, end to end: `pose_host` [3,E] (pinned CPU tensor of the engine's dtype) is copied in,
        and the kernel writes `ranges_host` [E,samples] (and `hits_host` int32) straight into PINNED host memory -- no device-to-host
        copy follows the launch; returns after the stream is synchronised."""
        t = self.torch
        for h in (pose_host, ranges_host) + (() if hits_host is None else (hits_host,)):
            if not (t.is_tensor(h) and h.is_pinned() and h.is_contiguous()):
                raise ValueError("scan_host needs contiguous pinned CPU tensors (tensor.pin_memory())")
        if ranges_host.dtype != self.engine.dtype or tuple(ranges_host.shape) != tuple(self.ranges.shape):
            raise ValueError("ranges_host must be [E, samples] of the engine's dtype")
        if self._pose_dev is None:
            self._pose_dev = t.empty((3, self.engine.E), dtype=self.engine.dtype, device=self.engine.device)
        self._pose_dev.copy_(pose_host, non_blocking=True)
        self._launch(self._pose_dev, ranges_host.data_ptr(), None if hits_host is None else hits_host.data_ptr())
        t.cuda.current_stream().synchronize()
        return ranges_host, hits_host


def scan_engine(engine, pose, range_, samples, max_distance, robot_radius=0.0, want_hits=True):
    """One-off device-resident scan (see EngineScanner): pose [3,E] tensor (x,y,yaw).  Returns (ranges, hits) device tensors."""
    pose = pose.to(device=engine.device, dtype=engine.dtype).contiguous()
    return EngineScanner(engine, range_, samples, max_distance, robot_radius, want_hits).scan(pose)
