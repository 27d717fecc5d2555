"""Single-env ``gym.Env`` views over a 1-env device handle.

Mirrors the public surface of the reference classes in
/root/reference/gym_reinmav/envs/native/ (``Quadrotor3D`` quadrotor3d.py:42-185 and twins):
``metadata``, ``action_space``, ``observation_space``, ``seed``, ``reset``, ``step``, ``render``,
``close`` and the public attributes ``state``, ``dt``, ``mass``, ``g``, ``ref_pos``, ``ref_vel``,
``pos_threshold``, ``vel_threshold`` (+ ``load_mass``, ``tether_length``).  ``step`` does NOT
reset on done — the caller does, as in test/test_quadrotor3d.py:21-22.

Every call goes through the CUDA path (one launch for one env: slow per step, but it is the
same kernel the batched VecEnv runs).  There is no NumPy implementation here.
"""
import os

import numpy as np

from ... import _capi
from ...batch import QuadBatch
from ...vec_env import make_spaces

try:                                    # be a real gym.Env when gym / gymnasium is installed
    import gym as _gym
    _EnvBase = _gym.Env
except Exception:                       # pragma: no cover - depends on the environment
    try:
        import gymnasium as _gym
        _EnvBase = _gym.Env
    except Exception:
        _EnvBase = object


class NativeQuadEnv(_EnvBase):
    metadata = {"render.modes": ["human"]}
    KIND = None
    REF_POS = None

    def __init__(self, device=None, seed=None):
        if device is None:
            device = int(os.environ.get("RMV_DEVICE", "0"))
        p = _capi.default_params(self.KIND)
        self.mass, self.dt = p.mass, p.dt
        planar = self.KIND in (_capi.QUAD2D, _capi.QUAD2D_SLUNG)
        if self.KIND == _capi.REINMAV:
            p.gravity = 9.81
        self.g = np.array([0.0, -p.gravity]) if planar else np.array([0.0, 0.0, -p.gravity])
        self.ref_pos = np.array(self.REF_POS, dtype=float)
        self.ref_vel = np.zeros_like(self.ref_pos)
        self.pos_threshold, self.vel_threshold = p.pos_done, p.vel_done
        if self.KIND in (_capi.QUAD2D_SLUNG, _capi.QUAD3D_SLUNG):
            self.load_mass, self.tether_length = p.load_mass, p.tether_len
        self.observation_space, self.action_space = make_spaces(self.KIND)
        self.viewer = None
        # viewer handles of the reference classes (quadrotor2d.py:57-59, quadrotor3d_slungload.py:62-72): never populated here
        # (render() is a no-op), present so that code reading them finds what the reference has before its first render()
        self.x_range = 1.0
        if planar:
            self.quadtrans = self.reftrans = None
            if self.KIND == _capi.QUAD2D_SLUNG:
                self.loadtrans = None
        elif self.KIND != _capi.REINMAV:
            for name in ("quad1", "quad2", "rotor1", "rotor2", "rotor3", "rotor4", "velocity", "ref"):
                setattr(self, "render_" + name, None)
            if self.KIND == _capi.QUAD3D_SLUNG:
                self.render_load = self.render_tether = None
        self._seed_value = self._fresh_seed() if seed is None else int(seed)
        self._batch = QuadBatch(self.KIND, 1, device=device, seed=self._seed_value)
        self._act = np.zeros((1, max(self._batch.action_dim, 1)), dtype=np.float32)
        self._rew = np.zeros(1, dtype=np.float32)
        self._done = np.zeros(1, dtype=np.uint8)
        self._obs = np.zeros((1, self._batch.state_dim), dtype=np.float32)
        self.reset()                    # the reference's __init__ seeds and resets (quadrotor3d.py:73-74)

    @staticmethod
    def _fresh_seed():
        return int.from_bytes(os.urandom(8), "little")

    # ------------------------------------------------------------------ gym.Env
    def seed(self, seed=None):
        self._seed_value = self._fresh_seed() if seed is None else int(seed)
        self._batch.set_seed(self._seed_value)
        return [self._seed_value]

    def reset(self):
        self._batch.reset_host(self._obs)
        return self._obs[0].astype(np.float64)

    def step(self, action):
        self._act[0, :] = np.asarray(action, dtype=np.float32).reshape(-1)
        self._batch.step_host(self._act, self._rew, self._done, self._obs, auto_reset=False)
        return self._obs[0].astype(np.float64), float(self._rew[0]), bool(self._done[0]), {}

    def render(self, mode="human", close=False):
        return None                     # viewers (pyglet / vpython) are outside this path

    def close(self):
        self._batch.close()

    # ------------------------------------------------------------------ reference attributes
    @property
    def state(self):
        return self._batch.get_state()[0].astype(np.float64)

    @state.setter
    def state(self, value):
        self._batch.set_state(np.asarray(value, dtype=np.float32).reshape(1, -1))

    @property
    def steps_beyond_done(self):
        """None until this env object has returned done once (quadrotor3d.py:68,114-122)."""
        return 0 if (int(self._batch.get_episode_words()[0]) & _capi.EVER_DONE_BIT) else None

    def control(self):
        """The reference's geometric controller for the current state (quadrotor3d.py:126-180 and
        twins), evaluated by the batched CUDA controller kernel; honours ``self.ref_pos`` / ``self.ref_vel``."""
        self._batch.control_host(self._act, ref_pos=self.ref_pos, ref_vel=self.ref_vel)
        return self._act[0].astype(np.float64)
