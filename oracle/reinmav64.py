"""reinmav-v0: vectorised float64 restatement of
/root/reference/gym_reinmav/envs/native/reinmav_env.py — the closed-loop 13-state rigid body
(position, velocity, quaternion, body rates) that tracks a quintic trajectory with its built-in PD
controller; ``step()`` takes NO action, returns reward 90 and done True (:111-116).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/reinmav-v0.npz, produced by
running the reference source (oracle/make_golden_reinmav.py).

One ``step()`` = explicit Euler over ``np.arange(t, t + dt, 1/5000)`` (:90-98): 50 or 51 substeps
depending on float rounding of the arange length, substep times ``t + i * ((t + ds) - t)`` (NumPy's
fill rule), then ``t += dt``.
"""
import numpy as np

ARM, MASS, GRAV = 0.0860, 0.1800, 9.8100          # :55-57
MIN_F, MAX_F = 0.0, 3.5316                         # :58-59
INERTIA = np.array([[0.00025, 0.0, 2.55e-06], [0.0, 0.000232, 0.0], [2.55e-06, 0.0, 0.0003738]])   # :60-62
INV_INERTIA = np.linalg.inv(INERTIA)                   # np.matrix.getI() of a square matrix                                           # :63
DT, DS = 1 / 100, 1 / 5000                         # :74, :91
INIT_STATE = np.array([0.0] * 6 + [1.0] + [0.0] * 6)                                             # :80
KP, KD = np.array([10.0, 10.0, 35.0]), np.array([5.0, 5.0, 22.0])                                 # :311-312
KP_ROT, KD_ROT = 100.0, 0.1                        # :313-314
K_QUAT = 2.0                                       # :242
EPS = np.finfo(np.float64).eps


def substep_times(t):
    """``np.arange(t, t + dt, ds)`` exactly as NumPy builds it."""
    n = int(np.ceil(((t + DT) - t) / DS))
    delta = (t + DS) - t
    return [t + i * delta for i in range(n)]


def trajectory(t):
    """trj_gen (:128-136): quintic 0 -> 1 in 4 s on x, y, z and yaw."""
    tau = np.maximum(0, np.minimum(t, 4.0)) / 4.0
    pos = 10.0 * tau ** 3 - 15.0 * tau ** 4 + 6.0 * tau ** 5
    vel = (30 / 4.0) * tau ** 2 - (60 / 4.0) * tau ** 3 + (30 / 4.0) * tau ** 4
    acc = (60 / 16.0) * tau - (180 / 16.0) * tau ** 2 + (120 / 16.0) * tau ** 3
    return pos, vel, acc


def quat_to_matrix(w, x, y, z):
    """quat2mat (:267-290)."""
    nq = w * w + x * x + y * y + z * z
    with np.errstate(all="ignore"):
        s = 2.0 / nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    m = [[1.0 - (yY + zZ), xY - wZ, xZ + wY],
         [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
         [xZ - wY, yZ + wX, 1.0 - (xX + yY)]]
    ok = nq > EPS
    eye = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    return [[np.where(ok, m[i][j], eye[i][j]) for j in range(3)] for i in range(3)]


def derivative(s, t):
    """quad_eq_of_motion1/2 + controller (:188-264, :306-337) for state ``s`` [13, N] at time ``t``."""
    pos, vel = s[0:3], s[3:6]
    qw, qx, qy, qz = s[6], s[7], s[8], s[9]
    p, q, r = s[10], s[11], s[12]
    R = quat_to_matrix(qw, qx, qy, qz)
    # RotToRPY (:341-346)
    phi = np.arcsin(R[1][2])
    cphi = np.cos(phi)
    psi = np.arctan2(-R[1][0] / cphi, R[1][1] / cphi)
    theta = np.arctan2(-R[0][2] / cphi, R[2][2] / cphi)
    d_pos, d_vel, d_acc = trajectory(t)
    psi_des, dpsi_des = d_pos, d_vel
    ddr = [d_acc + KD[i] * (d_vel - vel[i]) + KP[i] * (d_pos - pos[i]) for i in range(3)]
    u1 = MASS * (GRAV + ddr[2])
    sp, cp = np.sin(psi_des), np.cos(psi_des)
    phi_des = 1 / GRAV * (ddr[0] * sp - ddr[1] * cp)
    theta_des = 1 / GRAV * (ddr[0] * cp + ddr[1] * sp)
    mx = KP_ROT * (phi_des - phi) - KD_ROT * p
    my = KP_ROT * (theta_des - theta) - KD_ROT * q
    mz = KP_ROT * (psi_des - psi) + KD_ROT * (dpsi_des - r)
    # motor mixing with clamp (:204-216)
    k = 0.5 / ARM
    T = [0.25 * u1 - k * my, 0.25 * u1 + k * mx, 0.25 * u1 + k * my, 0.25 * u1 - k * mx]
    T = [np.maximum(np.minimum(ti, MAX_F / 4.0), MIN_F / 4.0) for ti in T]
    force = T[0] + T[1] + T[2] + T[3]
    mom = [ARM * T[1] - ARM * T[3], -ARM * T[0] + ARM * T[2], mz]
    # world acceleration: wRb = R^T, so wRb @ (0,0,F) = F * R[2][:]
    accel = [1.0 / MASS * (R[2][i] * force - (MASS * GRAV if i == 2 else 0.0)) for i in range(3)]
    err = 1 - (qw ** 2 + qx ** 2 + qy ** 2 + qz ** 2)
    qdot = [-0.5 * (-p * qx - q * qy - r * qz) + K_QUAT * err * qw,
            -0.5 * (p * qw - r * qy + q * qz) + K_QUAT * err * qx,
            -0.5 * (q * qw + r * qx - p * qz) + K_QUAT * err * qy,
            -0.5 * (r * qw - q * qx + p * qy) + K_QUAT * err * qz]
    om = [p, q, r]
    Iw = [INERTIA[i][0] * p + INERTIA[i][1] * q + INERTIA[i][2] * r for i in range(3)]
    cr = [om[1] * Iw[2] - om[2] * Iw[1], om[2] * Iw[0] - om[0] * Iw[2], om[0] * Iw[1] - om[1] * Iw[0]]
    rhs = [mom[i] - cr[i] for i in range(3)]
    pqr = [INV_INERTIA[i][0] * rhs[0] + INV_INERTIA[i][1] * rhs[1] + INV_INERTIA[i][2] * rhs[2] for i in range(3)]
    return np.stack([vel[0], vel[1], vel[2]] + accel + qdot + pqr)


class ReinmavBatch:
    def __init__(self, n_envs, state=None):
        self.n = n_envs
        self.state = np.tile(INIT_STATE[:, None], (1, n_envs)) if state is None else np.array(state, dtype=np.float64).T.copy()
        self.t = 0.0

    def step(self, round_to_float32=False):
        with np.errstate(all="ignore"):
            for ti in substep_times(self.t):
                self.state = self.state + DS * derivative(self.state, ti)
        if round_to_float32:                     # the CUDA path stores the state as float32 between steps
            self.state = self.state.astype(np.float32).astype(np.float64)
        self.t = self.t + DT
        return self.state.T.copy(), np.full(self.n, 90.0), np.ones(self.n, dtype=bool)
