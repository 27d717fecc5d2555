"""The four native env classes, same names and module tail (``...envs.native``) as the reference
(/root/reference/gym_reinmav/envs/native/__init__.py:2-5), so that ``gym_reinmav.run``'s
``env_type = entry_point.split(':')[0].split('.')[-1]`` (run.py:36,105) still yields 'native'."""
from ... import _capi
from .base import NativeQuadEnv


class Quadrotor2D(NativeQuadEnv):                 # quadrotor2d.py:41
    KIND = _capi.QUAD2D
    REF_POS = (0.0, 0.0)


class Quadrotor2DSlungload(NativeQuadEnv):        # quadrotor2d_slungload.py:41
    KIND = _capi.QUAD2D_SLUNG
    REF_POS = (0.0, 0.0)


class Quadrotor3D(NativeQuadEnv):                 # quadrotor3d.py:42
    KIND = _capi.QUAD3D
    REF_POS = (0.0, 0.0, 2.0)


class Quadrotor3DSlungload(NativeQuadEnv):        # quadrotor3d_slungload.py:42
    KIND = _capi.QUAD3D_SLUNG
    REF_POS = (0.0, 0.0, 1.0)


class ReinmavEnv(NativeQuadEnv):                  # reinmav_env.py:50
    """``reinmav-v0``: closed-loop rigid body with its own PD controller; ``step()`` takes no action and
    returns reward 90.0 / done True; ``reset()`` returns the current state (reinmav_env.py:99-126,348-351)."""
    KIND = _capi.REINMAV
    REF_POS = (0.0, 0.0, 0.0)

    def __init__(self, device=None, seed=None):
        super().__init__(device=device, seed=seed)
        self.arm_length, self.mass, self.gravity = 0.0860, 0.1800, 9.8100
        self.min_force, self.max_force = 0.0, 3.5316
        self.dt = 1 / 100

    @property
    def t(self):
        return self._batch.sim_time

    @t.setter
    def t(self, value):
        self._batch.sim_time = value

    def step(self, action=None):
        self._batch.step_host(None, self._rew, self._done, self._obs, auto_reset=False)
        return self._obs[0].astype(float), float(self._rew[0]), bool(self._done[0]), {}

    def control(self):
        raise TypeError("reinmav-v0 has no action input: its controller runs inside step()")
