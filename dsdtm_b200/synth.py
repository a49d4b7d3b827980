"""Synthetic RGB-D scenes for tests, smoke() and bench.py (SURVEY.md section 8d "synthetic inputs").

A textured height-field z = Z(X, Y) in world coordinates is ray-cast from a pinhole camera with pose
T_cw (p_cam = T_cw * p_world, the reference's ``mT_c2w``, ref: src/Frame.cpp:318-323). The texture is a sum
of low-frequency sinusoids plus two jittered lattices of Gaussian blobs so that FAST-10 fires at every
pyramid level and the photometric cost is smooth. Everything is float64 numpy and seeded, so the same
(seed, pose) always gives the same bytes. No oracle, no GPU code in here.
"""
import numpy as np

KINECT = dict(width=640, height=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, f=525.0)
# EuRoc.yaml intrinsics/size + Camera.f := fx
EUROC = dict(width=752, height=480, fx=458.654, fy=457.296, cx=367.215, cy=248.375, f=458.654)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    """splitmix64 finaliser on uint64 arrays."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def _u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class Scene:
    def __init__(self, seed=20260101):
        self.seed = np.uint64(seed)
        r = np.random.default_rng(int(seed))
        self.ph = r.uniform(0, 2 * np.pi, 6)
        self.z0 = 2.0
        self.za = 0.30
        # (cell size [m], sigma range [m], amplitude range, presence probability)
        self.lattices = [(0.05, (0.0045, 0.0085), (45.0, 100.0), 0.85),
                         (0.21, (0.018, 0.034), (35.0, 75.0), 0.9)]

    # ---- geometry
    def Z(self, X, Y):
        return self.z0 + self.za * np.sin(2.0 * X + self.ph[0]) * np.cos(1.5 * Y + self.ph[1])

    def gradZ(self, X, Y):
        sx, cx = np.sin(2.0 * X + self.ph[0]), np.cos(2.0 * X + self.ph[0])
        sy, cy = np.sin(1.5 * Y + self.ph[1]), np.cos(1.5 * Y + self.ph[1])
        return self.za * 2.0 * cx * cy, -self.za * 1.5 * sx * sy

    # ---- appearance
    def texture(self, X, Y):
        with np.errstate(over="ignore"):
            return self._texture(X, Y)

    def _texture(self, X, Y):
        ph = self.ph
        t = 120.0 + 28.0 * np.sin(3.0 * X + ph[2]) * np.cos(2.3 * Y + ph[3]) + 16.0 * np.sin(7.1 * X - 4.0 * Y + ph[4])
        for li, (c, (s0, s1), (a0, a1), prob) in enumerate(self.lattices):
            i0 = np.floor(X / c).astype(np.int64)
            j0 = np.floor(Y / c).astype(np.int64)
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    ii = i0 + di
                    jj = j0 + dj
                    key = (ii.astype(np.uint64) * np.uint64(0x100000001B3)) ^ (jj.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
                    key = key ^ (self.seed * np.uint64(0x632BE59BD9B4E019)) ^ np.uint64(li + 1)
                    h1 = _mix(key)
                    h2 = _mix(h1)
                    h3 = _mix(h2)
                    h4 = _mix(h3)
                    h5 = _mix(h4)
                    bx = (ii + 0.15 + 0.7 * _u01(h1)) * c
                    by = (jj + 0.15 + 0.7 * _u01(h2)) * c
                    sg = s0 + (s1 - s0) * _u01(h3)
                    u4 = _u01(h4)
                    amp = (a0 + (a1 - a0) * _u01(h5)) * np.where(u4 < 0.5 * prob, 1.0, np.where(u4 < prob, -1.0, 0.0))
                    t = t + amp * np.exp(-((X - bx) ** 2 + (Y - by) ** 2) / (2.0 * sg * sg))
        return t


# ---------------------------------------------------------------------------- SE3 (numpy, float64)
def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def pose_from_xi(xi):
    """[upsilon(3), omega(3)] -> pose7 {qw,qx,qy,qz,tx,ty,tz} (Sophus SE3::exp convention)."""
    xi = np.asarray(xi, np.float64)
    ups, om = xi[:3], xi[3:]
    th = np.linalg.norm(om)
    if th < 1e-12:
        q = np.array([1.0, *(0.5 * om)])
        V = np.eye(3)
    else:
        q = np.array([np.cos(th / 2), *(np.sin(th / 2) / th * om)])
        O = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * O + (th - np.sin(th)) / th ** 3 * (O @ O)
    q = q / np.linalg.norm(q)
    return np.concatenate([q, V @ ups])


def pose_mul(a, b):
    Ra, Rb = quat_to_R(a[:4]), quat_to_R(b[:4])
    aw, ax, ay, az = a[:4]
    bw, bx, by, bz = b[:4]
    q = np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                  aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx])
    q /= np.linalg.norm(q)
    return np.concatenate([q, a[4:] + Ra @ b[4:]])


def pose_inv(a):
    q = np.array([a[0], -a[1], -a[2], -a[3]])
    return np.concatenate([q, -(quat_to_R(q) @ a[4:])])


def pose_act(a, p):
    return (quat_to_R(a[:4]) @ np.asarray(p, np.float64).T).T + a[4:]


IDENTITY = np.array([1.0, 0, 0, 0, 0, 0, 0])


def pose_dist(a, b):
    """(rotation angle [rad], translation distance [m]) between two pose7."""
    d = pose_mul(pose_inv(a), b)
    ang = 2.0 * np.arctan2(np.linalg.norm(d[1:4]), abs(d[0]))
    return float(ang), float(np.linalg.norm(d[4:]))


# ---------------------------------------------------------------------------- rendering
def render(scene, cam, pose_cw, want_points=False):
    """Ray-cast the scene. Returns (u8 image HxW, depth HxW float64 = z in the camera frame [, world points HxWx3])."""
    w, h = cam["width"], cam["height"]
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    # float32 intrinsics, like the reference's Camera (ref: src/Camera.cpp:34-39)
    fx, fy, cx, cy = (float(np.float32(cam[k])) for k in ("fx", "fy", "cx", "cy"))
    dc = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)
    R = quat_to_R(pose_cw[:4])
    O = -(R.T @ pose_cw[4:])
    dw = dc @ R  # R^T d for every pixel
    s = (scene.z0 - O[2]) / dw[..., 2]
    for _ in range(8):
        X = O[0] + s * dw[..., 0]
        Y = O[1] + s * dw[..., 1]
        f = O[2] + s * dw[..., 2] - scene.Z(X, Y)
        gx, gy = scene.gradZ(X, Y)
        fp = dw[..., 2] - gx * dw[..., 0] - gy * dw[..., 1]
        s = s - f / fp
    X = O[0] + s * dw[..., 0]
    Y = O[1] + s * dw[..., 1]
    img = np.clip(np.rint(scene.texture(X, Y)), 0, 255).astype(np.uint8)
    if want_points:
        return img, s, np.stack([X, Y, O[2] + s * dw[..., 2]], -1)
    return img, s


def random_motion(rng, trans=0.02, rot_deg=0.5):
    xi = np.concatenate([rng.uniform(-trans, trans, 3), np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))])
    return pose_from_xi(xi)


def make_pair(seed, cam=KINECT, trans=0.02, rot_deg=0.5, ref_pose=None):
    """One synthetic frame pair. ref pose defaults to identity; cur = motion * ref (T_cur_ref = motion)."""
    rng = np.random.default_rng(seed)
    scene = Scene(seed)
    T_ref = IDENTITY.copy() if ref_pose is None else np.asarray(ref_pose, np.float64)
    T_c2r = random_motion(rng, trans, rot_deg)
    T_cur = pose_mul(T_c2r, T_ref)
    ref_img, ref_depth, ref_pts = render(scene, cam, T_ref, want_points=True)
    cur_img, _ = render(scene, cam, T_cur)
    return dict(scene=scene, cam=cam, T_ref=T_ref, T_cur=T_cur, T_c2r=T_c2r, ref_img=ref_img, cur_img=cur_img,
                ref_depth=ref_depth, ref_points=ref_pts)
