"""Vectorised CPU restatement of the four native ``step()`` / ``reset()`` pairs.

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

One code path, two arithmetic types:

* ``dtype=np.float64``  -> tier T0: what the reference computes (it works in
  float64 NumPy on 2-/3-vectors).  Pinned to the reference's own source by
  tests/golden/*.npz (made by oracle/make_golden.py from /root/reference).
* ``dtype=np.float32``  -> tier T1: the same formulas, same left-to-right
  evaluation order, rounded to float32; the ``acc + x*y`` steps the kernel
  issues as one FFMA go through ``_fma`` (an exact float32 fma, oracle/fma32.py;
  in float64 mode ``_fma`` is the plain two-rounding ``x*y + acc`` of the
  reference); sin/cos from oracle/sincos32.py.  The CUDA path must match T1
  bit for bit.

All paths below are relative to /root/reference/gym_reinmav/envs/native/.

State is struct-of-arrays ``[S, N]``.  Per-env bookkeeping the reference keeps
on the Python object: ``ever_done`` (``steps_beyond_done is not None``,
quadrotor3d.py:68,112-122 — set on the first done of an env instance and never
cleared by ``reset()``) and the episode counter that indexes the Philox reset
stream (replaces the MT19937 stream of quadrotor3d.py:77-79,184).
"""
import numpy as np

from . import philox
from .atan32 import atan2_32
from .fma32 import fma32
from .sincos32 import sincos32

QUAD2D, QUAD2D_SLUNG, QUAD3D, QUAD3D_SLUNG = 0, 1, 2, 3
KIND_OF_ID = {
    "quadrotor2d-v0": QUAD2D,
    "quadrotor2d-slungload-v0": QUAD2D_SLUNG,
    "quadrotor3d-v0": QUAD3D,
    "quadrotor3d-slungload-v0": QUAD3D_SLUNG,
}
STATE_DIM = {QUAD2D: 5, QUAD2D_SLUNG: 9, QUAD3D: 10, QUAD3D_SLUNG: 16}
ACTION_DIM = {QUAD2D: 2, QUAD2D_SLUNG: 2, QUAD3D: 4, QUAD3D_SLUNG: 4}
# U(action_space) bounds: quadrotor2d.py:62, quadrotor2d_slungload.py:68, quadrotor3d.py:70, quadrotor3d_slungload.py:75
ACTION_BOUNDS = {QUAD2D: (-10.0, 10.0), QUAD2D_SLUNG: (-10.0, 10.0), QUAD3D: (0.0, 10.0), QUAD3D_SLUNG: (-10.0, 10.0)}


def default_params(kind, repair="A"):
    """Constants hard-coded in each ``__init__`` (quadrotor3d.py:45-56,
    quadrotor3d_slungload.py:45-59, quadrotor2d.py:44-54 + literals :95-96,
    quadrotor2d_slungload.py:44-58)."""
    p = dict(mass=1.0, load_mass=0.1, dt=0.01, gravity=9.8, tether_len=0.0,
             pos_done=3.0, vel_done=10.0, thrust_scale=1.0, thrust_clamp=0,
             first_done_reward=1.0, later_done_reward=0.0,
             reward_scale=1.0)                               # run.py:89 reward_scale= -> baselines RewardScaler
    if kind == QUAD2D:
        p.update(thrust_scale=10.0, thrust_clamp=1)          # quadrotor2d.py:75-77
        if repair == "B":
            p.update(vel_done=2.0)                           # :54 via the repaired :96-98
    elif kind == QUAD2D_SLUNG:
        p.update(tether_len=0.5, pos_done=2.0)               # quadrotor2d_slungload.py:54-58
    elif kind == QUAD3D_SLUNG:
        p.update(tether_len=1.5)                             # quadrotor3d_slungload.py:59
    return p


class BatchOracle:
    def __init__(self, kind, n_envs, seed=0, env_id_offset=0, params=None, dtype=np.float64, repair="A"):
        if isinstance(kind, str):
            kind = KIND_OF_ID[kind]
        self.kind, self.n, self.seed = kind, int(n_envs), int(seed)
        self.dtype = np.dtype(dtype).type
        self.S, self.A = STATE_DIM[kind], ACTION_DIM[kind]
        self.env_ids = np.arange(env_id_offset, env_id_offset + self.n, dtype=np.uint64)
        self.p = default_params(kind, repair)
        if params:
            self.p.update(params)
        self.state = np.zeros((self.S, self.n), dtype=self.dtype)
        self.episode = np.zeros(self.n, dtype=np.uint32)      # number of resets so far
        self.ever_done = np.zeros(self.n, dtype=bool)
        self.taut_override = None           # tests only: force the slung-load branch
        self.tether_margin = None           # |tether| - L of the last step (slung kinds)
        self.done_margin = None             # min distance of the two done norms to their thresholds

    # ------------------------------------------------------------------ reset
    def _draws(self, sel):
        return philox.uniform_f32(self.seed, self.env_ids[sel], self.episode[sel],
                                  philox.STREAM_RESET_BASE + self.kind, self.S, -1.0, 1.0)

    def reset(self, mask=None):
        """``state = U[-1,1)^S`` (quadrotor3d.py:182-185 and twins); every call
        advances the episode counter of the envs it touches."""
        sel = np.arange(self.n) if mask is None else np.flatnonzero(mask)
        self.episode[sel] += np.uint32(1)
        self.state[:, sel] = self._draws(sel).astype(self.dtype)
        return self.obs()

    def obs(self):
        return np.ascontiguousarray(self.state.T)

    # ------------------------------------------------------------------ helpers
    def _c(self, v):
        return self.dtype(v)

    def _over(self, comps, n):
        """``vec / n``.  float64: the quotient, as the reference.  float32 (T1): one IEEE reciprocal
        and a product per component — what the CUDA path does (<= 1 ulp from the quotient)."""
        if self.dtype is np.float64:
            return [c / n for c in comps]
        inv = self._c(1.0) / n
        return [c * inv for c in comps]

    def _per_mass(self, x):
        """``x / self.mass`` (exact for the reference's mass = 1 either way)."""
        m = self._c(self.p["mass"])
        return x / m if self.dtype is np.float64 else x * (self._c(1.0) / m)

    def _fma(self, x, y, acc):
        """``x*y + acc``: two roundings in float64 (what NumPy does for the reference), ONE in float32."""
        if self.dtype is np.float64:
            return x * y + acc
        return fma32(x, y, acc)

    def _sumsq(self, comps):
        acc = comps[0] * comps[0]
        for c in comps[1:]:
            acc = self._fma(c, c, acc)
        return acc

    def _norm(self, comps):
        return np.sqrt(self._sumsq(comps))

    def _advance(self, x, v, a, velocity_first):
        """``x + v*dt + 0.5*a*dt*dt`` and ``v + a*dt`` (quadrotor3d.py:98-99; velocity first in
        quadrotor2d_slungload.py:100-101)."""
        dt, half = self._c(self.p["dt"]), self._c(0.5)
        if velocity_first:
            v2 = [self._fma(a[i], dt, v[i]) for i in range(len(x))]
            x2 = [self._fma(half * a[i] * dt, dt, self._fma(v2[i], dt, x[i])) for i in range(len(x))]
        else:
            x2 = [self._fma(half * a[i] * dt, dt, self._fma(v[i], dt, x[i])) for i in range(len(x))]
            v2 = [self._fma(a[i], dt, v[i]) for i in range(len(x))]
        return x2, v2

    def _thrust_axis_3d(self, q):
        """pyquaternion ``Quaternion(att).rotation_matrix.dot([0,0,1])``
        (quadrotor3d.py:94-96): normalise in place unless |1-q.q| < 1e-14, then
        column 3 of (Q Qbar^T)[1:,1:].  Returns (unit quaternion, R e3)."""
        w, x, y, z = q
        f = self._fma
        n2 = f(z, z, f(y, y, f(x, x, w * w)))
        n = np.sqrt(n2)
        need = ~(np.abs(self._c(1.0) - n2) < self._c(1e-14)) & (n > 0)
        safe = np.where(need, n, self._c(1.0))
        w, x, y, z = (np.where(need, c, q0) for c, q0 in zip(self._over((w, x, y, z), safe), (w, x, y, z)))
        rx = f(y, w, f(z, x, f(w, y, x * z)))
        ry = f(-x, w, f(-w, x, f(z, y, y * z)))
        rz = f(w, w, f(-x, x, f(-y, y, z * z)))
        return (w, x, y, z), (rx, ry, rz)

    def _qdot(self, qh, om):
        """``Quaternion.derivative(w)`` = (0.5*q) * (0, w) on the ALREADY
        normalised object (quadrotor3d.py:101 after :96)."""
        h = self._c(0.5)
        w, x, y, z = (h * c for c in qh)
        ox, oy, oz = om
        f = self._fma
        return (f(-z, oz, f(-y, oy, (-x) * ox)),
                f(y, oz, f(-z, oy, w * ox)),
                f(-x, oz, f(w, oy, z * ox)),
                f(w, oz, f(x, oy, (-y) * ox)))

    def _thrust(self, a0):
        """``thrust = 10.0 * action[0]; if thrust < 0: thrust = 0`` (quadrotor2d.py:75-77);
        scale 1 / no clamp for the other kinds (quadrotor3d.py:82)."""
        T = self._c(self.p["thrust_scale"]) * a0
        if self.p["thrust_clamp"]:
            T = np.where(T < 0, self._c(0.0), T)
        return T

    def _planar_axis(self, theta):
        """[cos(att+pi/2), sin(att+pi/2)] (quadrotor2d.py:88)."""
        if self.dtype is np.float64:
            a = theta + np.pi / 2
            return np.cos(a), np.sin(a)
        s, c = sincos32(theta)
        return -s, c

    def _taut(self, pos, lp, nt):
        """``linalg.norm(tether_vec) >= tether_length`` (quadrotor3d_slungload.py:104,
        quadrotor2d_slungload.py:96).

        After every taut step the constraint projection (:126-127) leaves the
        load at distance L(1 +- ulp), so this predicate sits on rounding noise.
        T1 therefore evaluates it in float64 FROM THE float32 STATE (exact
        differences, individually rounded float64 products/sums, IEEE sqrt): given
        the same float32 state the CUDA path and the float64 reference then take
        the same branch except on ~1e-16-relative ties.  ``taut_override`` lets a
        test force the branch to show that a fixture step is explained by one of
        the two."""
        L = self.p["tether_len"]
        if self.dtype is np.float64:
            margin = nt - L
        else:
            d = [lp[i].astype(np.float64) - pos[i].astype(np.float64) for i in range(len(pos))]
            sq = d[0] * d[0]
            for di in d[1:]:
                sq = sq + di * di                      # float64, individually rounded, as the kernel's rd ops
            margin = np.sqrt(sq) - np.float64(np.float32(L))
        self.tether_margin = margin
        if self.taut_override is not None:
            return np.asarray(self.taut_override, dtype=bool)
        return margin >= 0

    def _taut_branch(self, D, velocity_first, T, tm, axis, u, pos, vel, lp, lv, g):
        """Taut-tether dynamics + kinematic constraint (quadrotor3d_slungload.py:104-128,
        quadrotor2d_slungload.py:96-116).  Returns the unconstrained load / quad updates are folded in:
        (lp', lv', p', v', lp'', lv'')."""
        p, f = self.p, self._fma
        m, ml, L = self._c(p["mass"]), self._c(p["load_mass"]), self._c(p["tether_len"])
        thr = [T * axis[i] for i in range(D)]
        c = m * L * self._sumsq(lv)                         # SCALAR, broadcast-subtracted from every component (:110)
        inner = u[0] * (thr[0] - c)
        for i in range(1, D):
            inner = f(u[i], thr[i] - c, inner)
        inv = self._c(1.0) / (m + ml)
        la = [f(inv, inner * u[i], g[i]) for i in range(D)]
        lp_t, lv_t = self._advance(lp, lv, la, velocity_first)
        tn = ml * self._norm([-g[i] + la[i] for i in range(D)])          # the (x + g) - g rounding of :111,115
        inv_m = self._c(1.0) / m if self.dtype is not np.float64 else None
        if self.dtype is np.float64:
            acc = [f(tm, axis[i], g[i]) + (tn * u[i]) / m for i in range(D)]
        else:
            acc = [f(tn * u[i], inv_m, f(tm, axis[i], g[i])) for i in range(D)]
        p_t, v_t = self._advance(pos, vel, acc, velocity_first)
        dv = [lp_t[i] - p_t[i] for i in range(D)]
        d = self._over(dv, self._norm(dv))
        rel = (lv_t[0] - v_t[0]) * d[0]
        for i in range(1, D):
            rel = f(lv_t[i] - v_t[i], d[i], rel)
        lp_c = [f(d[i], L, p_t[i]) for i in range(D)]
        lv_c = [f(-rel, d[i], lv_t[i]) for i in range(D)]
        return lp_t, lv_t, p_t, v_t, lp_c, lv_c

    # ------------------------------------------------------------------ the four steps
    def _step_quad3d(self, s, a):
        """quadrotor3d.py:81-104."""
        p = self.p
        dt, m = self._c(p["dt"]), self._c(p["mass"])
        g = (self._c(0.0), self._c(0.0), self._c(-p["gravity"]))
        T, om = self._thrust(a[0]), (a[1], a[2], a[3])
        pos, q, vel = s[0:3], s[3:7], s[7:10]
        qh, e3 = self._thrust_axis_3d(q)
        tm = self._per_mass(T)
        acc = [self._fma(tm, e3[i], g[i]) for i in range(3)]
        pos2, vel2 = self._advance(pos, vel, acc, False)
        qd = self._qdot(qh, om)
        q2 = [self._fma(qd[i], dt, q[i]) for i in range(4)]
        new = pos2 + q2 + vel2
        return new, self._norm(pos2), self._norm(vel2), self._norm(pos2)

    def _step_quad3d_slung(self, s, a):
        """quadrotor3d_slungload.py:87-147."""
        p = self.p
        dt, m, ml, L = self._c(p["dt"]), self._c(p["mass"]), self._c(p["load_mass"]), self._c(p["tether_len"])
        g = (self._c(0.0), self._c(0.0), self._c(-p["gravity"]))
        half = self._c(0.5)
        T, om = self._thrust(a[0]), (a[1], a[2], a[3])
        pos, q, vel, lp, lv = s[0:3], s[3:7], s[7:10], s[10:13], s[13:16]
        tv = [lp[i] - pos[i] for i in range(3)]
        nt = self._norm(tv)
        u = self._over(tv, nt)
        taut = self._taut(pos, lp, nt)
        qh, e3 = self._thrust_axis_3d(q)
        qd = self._qdot(qh, om)
        q2 = [self._fma(qd[i], dt, q[i]) for i in range(4)]
        tm = self._per_mass(T)
        lp_t, lv_t, p_t, v_t, lp_c, lv_c = self._taut_branch(3, False, T, tm, e3, u, pos, vel, lp, lv, g)   # (:104-128)
        # --- slack (:131-145)
        lp_s, lv_s = self._advance(lp, lv, g, False)
        acc_s = [self._fma(tm, e3[i], g[i]) for i in range(3)]
        p_s, v_s = self._advance(pos, vel, acc_s, False)
        sel = lambda t_, s_: [np.where(taut, t_[i], s_[i]) for i in range(len(t_))]
        pos2, vel2, lp2, lv2 = sel(p_t, p_s), sel(v_t, v_s), sel(lp_c, lp_s), sel(lv_c, lv_s)
        new = pos2 + q2 + vel2 + lp2 + lv2
        # done: LOAD position, QUAD velocity (:149-152); reward: load position (:156)
        return new, self._norm(lp2), self._norm(vel2), self._norm(lp2)

    def _step_quad2d(self, s, a):
        """quadrotor2d.py:74-93 (+ done literals :95-96)."""
        p = self.p
        dt, m, half = self._c(p["dt"]), self._c(p["mass"]), self._c(0.5)
        g = (self._c(0.0), self._c(-p["gravity"]))
        T, om = self._thrust(a[0]), a[1]
        pos, th, vel = s[0:2], s[2], s[3:5]
        ax = self._planar_axis(th)
        tm = self._per_mass(T)
        acc = [self._fma(tm, ax[i], g[i]) for i in range(2)]
        pos2, vel2 = self._advance(pos, vel, acc, False)
        th2 = self._fma(om, dt, th)
        new = pos2 + [th2] + vel2
        return new, self._norm(pos2), self._norm(vel2), self._norm(pos2)

    def _step_quad2d_slung(self, s, a):
        """quadrotor2d_slungload.py:79-134."""
        p = self.p
        dt, m, ml, L = self._c(p["dt"]), self._c(p["mass"]), self._c(p["load_mass"]), self._c(p["tether_len"])
        g = (self._c(0.0), self._c(-p["gravity"]))
        half = self._c(0.5)
        T, om = self._thrust(a[0]), a[1]
        pos, th, vel, lp, lv = s[0:2], s[2], s[3:5], s[5:7], s[7:9]
        tv = [lp[i] - pos[i] for i in range(2)]
        nt = self._norm(tv)
        u = self._over(tv, nt)
        taut = self._taut(pos, lp, nt)
        ax = self._planar_axis(th)
        tm = self._per_mass(T)
        th2 = self._fma(om, dt, th)
        # taut (:96-116): velocity first, then position with the NEW velocity
        lp_t, lv_t, p_t, v_t, lp_c, lv_c = self._taut_branch(2, True, T, tm, ax, u, pos, vel, lp, lv, g)
        # --- slack (:119-132)
        lp_s, lv_s = self._advance(lp, lv, g, True)
        acc_s = [self._fma(tm, ax[i], g[i]) for i in range(2)]
        p_s, v_s = self._advance(pos, vel, acc_s, True)
        sel = lambda t_, s_: [np.where(taut, t_[i], s_[i]) for i in range(len(t_))]
        pos2, vel2, lp2, lv2 = sel(p_t, p_s), sel(v_t, v_s), sel(lp_c, lp_s), sel(lv_c, lv_s)
        new = pos2 + [th2] + vel2 + lp2 + lv2
        # done: load position / load velocity (:136-139); reward: QUAD position (:143)
        return new, self._norm(lp2), self._norm(lv2), self._norm(pos2)

    # ------------------------------------------------------------------ control()
    REF_POS = {QUAD2D: (0.0, 0.0), QUAD2D_SLUNG: (0.0, 0.0), QUAD3D: (0.0, 0.0, 2.0), QUAD3D_SLUNG: (0.0, 0.0, 1.0)}

    def control(self, ref_pos=None, ref_vel=None):
        """The reference's geometric controllers, vectorised: ``[N, A]`` actions from the current state.

        2D kinds: quadrotor2d.py:115-138 / quadrotor2d_slungload.py:156-183 (PD acceleration, attitude
        from atan2, first-order attitude loop tau = 0.1).  3D kinds: quadrotor3d.py:126-180 /
        quadrotor3d_slungload.py:169-226 (PD acceleration, desired attitude via ``Quaternion(matrix=)``
        — pyquaternion's trace method on the transposed matrix — error quaternion from the
        UN-normalised current attitude, tau = 0.3, thrust projected on the normalised body z).

        float32 (tier T1) is the bit model of ``control_law`` in csrc/rmv_math.cuh: the same expression order,
        every operation individually rounded (NumPy float32 arithmetic), the sum of squares inside the norms and the
        body-z projection fused as in ``step()``, ``atan2`` from oracle/atan32.py."""
        s = self.state
        rp = [self._c(v) for v in (ref_pos if ref_pos is not None else self.REF_POS[self.kind])]
        rv = [self._c(v) for v in (ref_vel if ref_vel is not None else (0.0,) * len(rp))]
        gr = self._c(self.p["gravity"])
        with np.errstate(all="ignore"):
            if self.kind in (QUAD2D, QUAD2D_SLUNG):
                kp, kv, tau = self._c(-5.0), self._c(-4.0), self._c(0.1)
                pos, th, vel = s[0:2], s[2], s[3:5]
                acc = [kp * (pos[i] - rp[i]) + kv * (vel[i] - rv[i]) + (gr if i == 1 else self._c(0.0)) for i in range(2)]
                atan2 = np.arctan2 if self.dtype is np.float64 else atan2_32
                att = atan2(acc[1], acc[0]) - self._c(np.pi / 2)
                w = (self._c(-1.0) / tau) * (th - att)
                thrust = self._c(self.p["mass"]) * self._norm(acc)
                return np.stack([thrust, w], axis=1)
            kp, kv, tau = self._c(-5.0), self._c(-4.0), self._c(0.3)
            pos, q, vel = s[0:3], s[3:7], s[7:10]
            acc = [kp * (pos[i] - rp[i]) + kv * (vel[i] - rv[i]) - (-gr if i == 2 else self._c(0.0)) for i in range(3)]
            na = self._norm(acc)
            zb = [a / na for a in acc]
            # xb = yc x zb with yc = (0,1,0); yb = zb x xb
            xb = [zb[2], 0 * zb[0], -zb[0]]
            nx = self._norm(xb)
            xb = [x / nx for x in xb]
            yb = [zb[1] * xb[2] - zb[2] * xb[1], zb[2] * xb[0] - zb[0] * xb[2], zb[0] * xb[1] - zb[1] * xb[0]]
            nz = self._norm(zb)
            zb = [z / nz for z in zb]
            # rotmat columns are (xb, yb, zb); pyquaternion works on m = rotmat^T, i.e. rows (xb, yb, zb)
            m = [xb, yb, zb]
            qd = self._quat_from_rows(m)
            # error = conj(q_current) * q_desired with the raw (un-normalised) state quaternion
            w0, x0, y0, z0 = q[0], -q[1], -q[2], -q[3]
            dw, dx, dy, dz = qd
            ew = w0 * dw - x0 * dx - y0 * dy - z0 * dz
            ex = x0 * dw + w0 * dx - z0 * dy + y0 * dz
            ey = y0 * dw + z0 * dx + w0 * dy - x0 * dz
            ez = z0 * dw - y0 * dx + x0 * dy + w0 * dz
            gain = (self._c(2.0) / tau) * np.sign(ew)
            _, e3 = self._thrust_axis_3d(q)
            thrust = acc[0] * e3[0] + acc[1] * e3[1] + acc[2] * e3[2]
            return np.stack([thrust, gain * ex, gain * ey, gain * ez], axis=1)

    def _quat_from_rows(self, m):
        """pyquaternion ``Quaternion(matrix=R)`` after its ``m = R^T``: the
        branch on m22 / m00 vs m11 picks the numerically largest component."""
        one = self._c(1.0)
        m00, m01, m02 = m[0]
        m10, m11, m12 = m[1]
        m20, m21, m22 = m[2]
        c1 = m22 < 0
        c2 = m00 > m11
        c3 = m00 < -m11
        ta, qa = one + m00 - m11 - m22, (m12 - m21, None, m01 + m10, m20 + m02)
        tb, qb = one - m00 + m11 - m22, (m20 - m02, m01 + m10, None, m12 + m21)
        tc, qc = one - m00 - m11 + m22, (m01 - m10, m20 + m02, m12 + m21, None)
        td, qd = one + m00 + m11 + m22, (None, m12 - m21, m20 - m02, m01 - m10)
        fill = lambda qq, t: [t if c is None else c for c in qq]
        qa, qb, qc, qd = fill(qa, ta), fill(qb, tb), fill(qc, tc), fill(qd, td)
        t = np.where(c1, np.where(c2, ta, tb), np.where(c3, tc, td))
        comps = [np.where(c1, np.where(c2, qa[i], qb[i]), np.where(c3, qc[i], qd[i])) for i in range(4)]
        scale = self._c(0.5) / np.sqrt(t)
        return [c * scale for c in comps]

    # ------------------------------------------------------------------ public step
    def step(self, actions, auto_reset=True):
        """One step of all envs.  ``actions`` is ``[N, A]``.

        Returns (obs [N,S], reward [N], done [N] bool).  With ``auto_reset`` the
        done envs are re-drawn and ``obs`` carries the reset state (the VecEnv
        convention the reference's launcher relies on, gym_reinmav/run.py:89);
        without it the terminal state stays (gym.Env convention,
        test/test_quadrotor3d.py:21-22)."""
        a = np.asarray(actions).astype(self.dtype).T
        fn = {QUAD2D: self._step_quad2d, QUAD2D_SLUNG: self._step_quad2d_slung,
              QUAD3D: self._step_quad3d, QUAD3D_SLUNG: self._step_quad3d_slung}[self.kind]
        with np.errstate(all="ignore"):
            new, n_pos, n_vel, n_rew = fn(self.state, a)
            # strict '>' on the norms; the '< -threshold' clauses are dead (quadrotor3d.py:106-110); NaN -> not done
            done = (n_pos > self._c(self.p["pos_done"])) | (n_vel > self._c(self.p["vel_done"]))
            self.done_margin = np.minimum(np.abs(n_pos - self._c(self.p["pos_done"])),
                                          np.abs(n_vel - self._c(self.p["vel_done"])))
        self.state = np.stack([np.asarray(c, dtype=self.dtype) for c in new])
        # reward (quadrotor3d.py:112-122)
        reward = np.where(done,
                          np.where(self.ever_done, self._c(self.p["later_done_reward"]), self._c(self.p["first_done_reward"])),
                          -n_rew).astype(self.dtype)
        self.ever_done |= done
        self.raw_reward = reward                           # what Monitor accumulates (RewardScaler wraps Monitor)
        if self.p["reward_scale"] != 1.0:
            reward = (reward * self._c(self.p["reward_scale"])).astype(self.dtype)
        if auto_reset and done.any():
            self.reset(mask=done)
        return self.obs(), reward, done
