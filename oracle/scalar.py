"""T0s: one-object-per-env float64 port with the reference's execution model.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the CPU arm that
``bench.py`` times (``cpu_baseline`` and ``--impl reference``): like the
reference it keeps one Python object per env and does its arithmetic with
NumPy calls on 2-/3-/4-element float64 arrays — ``np.linalg.norm``,
``np.inner``, small ``np.dot`` — so the cost profile (interpreter + tiny-array
overhead, ~40 NumPy calls per 3D step) and the rounding of every reduction are
the reference's.  On the machine that generated tests/golden it reproduces the
reference trajectories bit for bit (tests/test_oracle_golden.py), including
through the slung-load taut/slack switch, which sits on 1-ulp noise after every
taut step (DESIGN.md "Switching surface").

Formulas follow /root/reference/gym_reinmav/envs/native/:
  quadrotor2d.py:74-113, quadrotor2d_slungload.py:79-154,
  quadrotor3d.py:81-124, quadrotor3d_slungload.py:87-167 and the ``reset()``s.
"""
import math

import numpy as np
from numpy.linalg import norm

from . import philox
from .batch import (ACTION_DIM, KIND_OF_ID, QUAD2D, QUAD2D_SLUNG, QUAD3D, QUAD3D_SLUNG,
                    STATE_DIM, default_params)

_E3 = np.array([0.0, 0.0, 1.0])
_HALF_PI = np.pi / 2


def _unit_and_axis(att):
    """pyquaternion semantics (quadrotor3d.py:94-96): returns the in-place
    normalised quaternion and rotation_matrix.dot(e3)."""
    q = np.array(att, dtype=float)
    sos = np.dot(q, q)
    if not abs(1.0 - sos) < 1e-14:
        n = np.sqrt(sos)
        if n > 0:
            q = q / n
    w, x, y, z = q
    left = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
    right = np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])
    rot = np.dot(left, right.conj().transpose())[1:][:, 1:]
    return q, left, rot.dot(_E3)


def _rate(q_unit, omega):
    """Quaternion.derivative: (0.5*q) (x) (0, omega) (quadrotor3d.py:101)."""
    hw, hx, hy, hz = np.dot(np.array([[0.5, -0.0, -0.0, -0.0], [0.0, 0.5, -0.0, 0.0],
                                       [0.0, 0.0, 0.5, -0.0], [0.0, -0.0, 0.0, 0.5]]), q_unit)
    m = np.array([[hw, -hx, -hy, -hz], [hx, hw, -hz, hy], [hy, hz, hw, -hx], [hz, -hy, hx, hw]])
    return np.dot(m, np.array([0.0, omega[0], omega[1], omega[2]]))


class ScalarEnv:
    """A single env; ``step`` does not auto-reset (gym.Env convention)."""

    def __init__(self, kind, seed=0, env_id=0, params=None, repair="A"):
        if isinstance(kind, str):
            kind = KIND_OF_ID[kind]
        self.kind, self.seed, self.env_id = kind, seed, env_id
        self.S, self.A = STATE_DIM[kind], ACTION_DIM[kind]
        self.p = default_params(kind, repair)
        if params:
            self.p.update(params)
        self.planar = kind in (QUAD2D, QUAD2D_SLUNG)
        self.slung = kind in (QUAD2D_SLUNG, QUAD3D_SLUNG)
        d = 2 if self.planar else 3
        self.g = np.zeros(d)
        self.g[-1] = -self.p["gravity"]
        self.episode = 0
        self.ever_done = False
        self.state = np.zeros(self.S)

    def reset(self):
        self.episode += 1
        self.state = philox.uniform_f32(self.seed, [self.env_id], [self.episode],
                                        philox.STREAM_RESET_BASE + self.kind, self.S, -1.0, 1.0)[:, 0].astype(np.float64)
        return self.state.copy()

    def _split(self):
        s = self.state
        if self.planar:
            parts = [s[0:2], s[2:3], s[3:5]] + ([s[5:7], s[7:9]] if self.slung else [])
        else:
            parts = [s[0:3], s[3:7], s[7:10]] + ([s[10:13], s[13:16]] if self.slung else [])
        return [np.array(v, dtype=float) for v in parts]

    def step(self, action):
        p = self.p
        dt, m = p["dt"], p["mass"]
        thrust = p["thrust_scale"] * action[0] if p["thrust_scale"] != 1.0 else action[0]
        if p["thrust_clamp"] and thrust < 0.0:
            thrust = 0.0
        parts = self._split()
        pos, att, vel = parts[0], parts[1], parts[2]
        if self.planar:
            omega = action[1]
            ang = att + _HALF_PI
            axis = np.array([math.cos(ang[0]), math.sin(ang[0])])
            att_rate = omega
        else:
            omega = action[1:4]
            q_unit, _, axis = _unit_and_axis(att)
            att_rate = _rate(q_unit, omega)
        g = self.g
        velocity_first = self.kind == QUAD2D_SLUNG            # quadrotor2d_slungload.py:100-101,109-110

        def advance(x, v, a):
            if velocity_first:
                v2 = v + a * dt
                return x + v2 * dt + 0.5 * a * dt * dt, v2
            return x + v * dt + 0.5 * a * dt * dt, v + a * dt

        if not self.slung:
            acc = thrust / m * axis + g
            pos, vel = advance(pos, vel, acc)
            att = att + att_rate * dt
            new = [pos, att, vel]
            n_pos, n_vel, n_rew = norm(pos, 2), norm(vel, 2), norm(pos, 2)
        else:
            lpos, lvel = parts[3], parts[4]
            ml, length = p["load_mass"], p["tether_len"]
            tether = lpos - pos
            unit = tether / norm(tether)
            if norm(tether) >= length:
                pull = thrust * axis
                la = np.inner(unit, pull - m * length * np.inner(lvel, lvel)) * unit
                la = (1 / (m + ml)) * la + g
                lpos, lvel = advance(lpos, lvel, la)
                tension = ml * norm(-g + la) * unit
                acc = thrust / m * axis + g + tension / m
                pos, vel = advance(pos, vel, acc)
                att = att + att_rate * dt
                direction = (lpos - pos) / norm(lpos - pos)
                lpos = pos + direction * length
                lvel = lvel - np.inner(lvel - vel, direction) * direction
            else:
                lpos, lvel = advance(lpos, lvel, g)
                acc = thrust / m * axis + g
                pos, vel = advance(pos, vel, acc)
                att = att + att_rate * dt
            new = [pos, att, vel, lpos, lvel]
            if self.kind == QUAD2D_SLUNG:                        # :136-143
                n_pos, n_vel, n_rew = norm(lpos, 2), norm(lvel, 2), norm(pos, 2)
            else:                                                # quadrotor3d_slungload.py:149-156
                n_pos, n_vel, n_rew = norm(lpos, 2), norm(vel, 2), norm(lpos, 2)
        self.state = np.concatenate(new)
        done = bool(n_pos > p["pos_done"] or n_vel > p["vel_done"])
        if not done:
            reward = -n_rew
        elif not self.ever_done:
            self.ever_done = True
            reward = p["first_done_reward"]
        else:
            reward = p["later_done_reward"]
        return self.state.copy(), float(reward), done


def rollout(kind, n_steps, seed=0, env_id=0, actions=None):
    """Random-policy loop with reset-on-done; returns (#steps, #dones)."""
    from .batch import ACTION_BOUNDS
    env = ScalarEnv(kind, seed=seed, env_id=env_id)
    env.reset()
    lo, hi = ACTION_BOUNDS[env.kind]
    if actions is None:
        actions = np.random.RandomState(seed + 7919 * env_id).uniform(lo, hi, size=(n_steps, env.A))
    dones = 0
    for t in range(n_steps):
        _, _, d = env.step(actions[t])
        if d:
            dones += 1
            env.reset()
    return n_steps, dones
