"""Synthetic HDL-64E-like scans for the benchmark configs of BASELINE.json.

Analytic ray casting of 64 beams x 2048 azimuth steps (= 131 072 rays) into a closed scene: ground plane
z = -1.73 m, a 120 x 120 x 15 m room (so every ray returns), axis-aligned boxes and vertical cylinders.
Pure numpy, deterministic per seed; host-side input generation only (not part of the timed path).
"""
import numpy as np

N_BEAMS = 64
N_AZIMUTH = 2048
GROUND_Z = -1.73
ROOM_HALF = 60.0
ROOM_HEIGHT = 15.0


class Scene:
    def __init__(self, seed=20260117, n_boxes=40, n_cylinders=20):
        rng = np.random.default_rng(seed)
        # boxes: centre in an annulus around the origin so the sensor path (|xy| < 4 m) stays free
        r = rng.uniform(7.0, 55.0, n_boxes)
        a = rng.uniform(0, 2 * np.pi, n_boxes)
        size = rng.uniform(1.0, 8.0, (n_boxes, 3))
        cx, cy = r * np.cos(a), r * np.sin(a)
        self.box_lo = np.stack([cx - size[:, 0] / 2, cy - size[:, 1] / 2, np.full(n_boxes, GROUND_Z)], 1)
        self.box_hi = np.stack([cx + size[:, 0] / 2, cy + size[:, 1] / 2, GROUND_Z + size[:, 2]], 1)
        r = rng.uniform(5.0, 50.0, n_cylinders)
        a = rng.uniform(0, 2 * np.pi, n_cylinders)
        self.cyl_xy = np.stack([r * np.cos(a), r * np.sin(a)], 1)
        self.cyl_r = rng.uniform(0.2, 0.6, n_cylinders)
        self.cyl_top = GROUND_Z + rng.uniform(3.0, 10.0, n_cylinders)

    def cast(self, origin, dirs):
        """Distance along unit `dirs` (n,3) from `origin` (3,) to the first surface."""
        o = np.asarray(origin, np.float64)
        d = np.asarray(dirs, np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            # room (inside-out box): exit distance
            lo = np.array([-ROOM_HALF, -ROOM_HALF, GROUND_Z])
            hi = np.array([ROOM_HALF, ROOM_HALF, GROUND_Z + ROOM_HEIGHT])
            t_exit = np.where(d > 0, (hi - o) * inv, np.where(d < 0, (lo - o) * inv, np.inf))
            best = t_exit.min(axis=1)
            for blo, bhi in zip(self.box_lo, self.box_hi):
                t0 = (blo - o) * inv
                t1 = (bhi - o) * inv
                tn = np.minimum(t0, t1)
                tf = np.maximum(t0, t1)
                tn = np.where(np.isnan(tn), -np.inf, tn).max(axis=1)
                tf = np.where(np.isnan(tf), np.inf, tf).min(axis=1)
                hit = (tn <= tf) & (tn > 0)
                best = np.where(hit & (tn < best), tn, best)
            for (cx, cy), cr, ctop in zip(self.cyl_xy, self.cyl_r, self.cyl_top):
                ox, oy = o[0] - cx, o[1] - cy
                a = d[:, 0] ** 2 + d[:, 1] ** 2
                b = 2 * (ox * d[:, 0] + oy * d[:, 1])
                c = ox * ox + oy * oy - cr * cr
                disc = b * b - 4 * a * c
                t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * a)
                z = o[2] + t * d[:, 2]
                hit = (disc > 0) & (t > 0) & (z <= ctop) & (z >= GROUND_Z)
                best = np.where(hit & (t < best), t, best)
        return best


def beam_directions():
    """(131072, 3) unit vectors in the sensor frame, beam-major like a spinning 64-beam lidar."""
    elev = np.deg2rad(np.linspace(2.0, -24.8, N_BEAMS))
    azim = np.linspace(0.0, 2 * np.pi, N_AZIMUTH, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(azim)[None, :], ce * np.sin(azim)[None, :], np.broadcast_to(se, (N_BEAMS, N_AZIMUTH))], -1)
    return d.reshape(-1, 3)


def rot_zyx(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def make_pose(t, rpy):
    T = np.eye(4)
    T[:3, :3] = rot_zyx(*rpy)
    T[:3, 3] = t
    return T


def scan(scene, pose, rng, noise=0.02, azimuth_stride=1):
    """One scan taken at `pose` (4x4 sensor->world). Returns (131072 / azimuth_stride, 3) float32 points in the sensor frame."""
    d_local = beam_directions()
    if azimuth_stride > 1:
        d_local = d_local.reshape(N_BEAMS, N_AZIMUTH, 3)[:, ::azimuth_stride].reshape(-1, 3)
    d_world = d_local @ pose[:3, :3].T
    rng_m = scene.cast(pose[:3, 3], d_world)
    rng_m = rng_m + rng.normal(0.0, noise, rng_m.shape)
    return (d_local * rng_m[:, None]).astype(np.float32)


def ranges(scene, pose):
    """Noise-free ranges of the 131072 beams at `pose` (for trajectories that revisit poses: cast once, add noise per visit)."""
    return scene.cast(pose[:3, 3], beam_directions() @ pose[:3, :3].T).astype(np.float32)


def scan_from_ranges(rng_m, rng, noise=0.02):
    return (beam_directions() * (rng_m.astype(np.float64) + rng.normal(0.0, noise, rng_m.shape))[:, None]).astype(np.float32)


def figure_eight_poses(n, radius=1.5, step=0.7, lateral=0.0):
    """closed figure-eight (two tangent circles of `radius` around the origin, inside the scene's free zone),
    `step` metres per frame, heading along the path; `lateral` shifts the whole lap sideways (a second visit of a place)."""
    poses = []
    for i in range(n):
        u = (i * step / radius) % (4 * np.pi)
        if u < 2 * np.pi:
            x, y, yaw = radius * np.sin(u), radius * (1 - np.cos(u)), u
        else:
            v = u - 2 * np.pi
            x, y, yaw = radius * np.sin(v), -radius * (1 - np.cos(v)), -v
        # sideways offset along the path normal
        poses.append(make_pose([x - lateral * np.sin(yaw), y + lateral * np.cos(yaw), 0.0], [0.0, 0.0, yaw]))
    return poses


def odometry_motion(rng):
    """Relative motion of SURVEY C2: tx~U(0.4,1.0), ty~U(-.05,.05), tz~U(-.02,.02), yaw~U(-2,2) deg, roll/pitch~U(-.3,.3) deg."""
    t = [rng.uniform(0.4, 1.0), rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02)]
    rpy = np.deg2rad([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-2.0, 2.0)])
    return make_pose(t, rpy)


def loop_motion(rng):
    """Relative motion of SURVEY C4: tx,ty~U(-2,2), yaw~U(-10,10) deg."""
    t = [rng.uniform(-2.0, 2.0), rng.uniform(-2.0, 2.0), rng.uniform(-0.02, 0.02)]
    rpy = np.deg2rad([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-10.0, 10.0)])
    return make_pose(t, rpy)


def scan_pair(seed=20260117, loop=False, scene_seed=None):
    """(source, target, truth): two scans of one scene; truth = pose of target in source frame (what align returns)."""
    rng = np.random.default_rng(seed)
    scene = Scene(seed if scene_seed is None else scene_seed)
    pose0 = np.eye(4)
    rel = loop_motion(rng) if loop else odometry_motion(rng)
    pose1 = pose0 @ rel
    return scan(scene, pose0, rng), scan(scene, pose1, rng), rel


def map_cloud(seed=20260117, n_scans=16, path_len=12.0):
    """n_scans scans along a straight path, concatenated in the first scan's frame (16 -> 2 097 152 points)."""
    rng = np.random.default_rng(seed)
    scene = Scene(seed)
    out = []
    for i in range(n_scans):
        pose = make_pose([path_len * i / max(n_scans - 1, 1) - path_len / 2, 0.0, 0.0], [0.0, 0.0, 0.0])
        p = scan(scene, pose, rng).astype(np.float64)
        out.append((p @ pose[:3, :3].T + pose[:3, 3]).astype(np.float32))
    return np.concatenate(out, 0)


def trajectory(seed=20260117, n_scans=9):
    """n_scans consecutive scans of one scene along an odometry path; returns (scans, poses) with poses[i] = scan i in the world."""
    rng = np.random.default_rng(seed)
    scene = Scene(seed)
    pose = np.eye(4)
    scans, poses = [], []
    for i in range(n_scans):
        scans.append(scan(scene, pose, rng))
        poses.append(pose.copy())
        pose = pose @ odometry_motion(rng)
    return scans, poses
