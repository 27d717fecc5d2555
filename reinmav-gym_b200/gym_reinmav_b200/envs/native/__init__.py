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
