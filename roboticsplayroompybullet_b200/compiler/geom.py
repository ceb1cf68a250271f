"""Small fp64 rigid-transform helpers used by the host-side model compiler.

Conventions follow the reference's physics backend (Bullet): quaternions are
[x, y, z, w]; Euler angles are URDF roll-pitch-yaw, R = Rz(yaw) Ry(pitch) Rx(roll)
(what `getQuaternionFromEuler` returns, reference call site environments.py:960).
"""
import numpy as np


def rpy_to_mat(rpy):
    r, p, y = [float(v) for v in rpy]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]])


def rpy_to_quat(rpy):
    r, p, y = [0.5 * float(v) for v in rpy]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def quat_to_mat(q):
    x, y, z, w = [float(v) for v in q]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([
        [1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
        [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
        [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def mat_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0, 0, 0, 0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return np.array(q)


class Xf:
    """Rigid transform: x_parent = R @ x_child + p."""

    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64)
        self.p = np.zeros(3) if p is None else np.array(p, dtype=np.float64)

    def __mul__(self, o):
        return Xf(self.R @ o.R, self.R @ o.p + self.p)

    def inv(self):
        return Xf(self.R.T, -self.R.T @ self.p)

    def apply(self, v):
        return self.R @ np.asarray(v, dtype=np.float64) + self.p


def box_inertia(mass, full_dims):
    lx, ly, lz = full_dims
    return mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])


def parallel_axis(m, d):
    d = np.asarray(d, dtype=np.float64)
    return m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
