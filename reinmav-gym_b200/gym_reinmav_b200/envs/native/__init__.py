"""The four native env classes, same names and module tail (``...envs.native``) as the reference
(/root/reference/gym_reinmav/envs/native/__init__.py:2-5), so that ``gym_reinmav.run``'s
``env_type = entry_point.split(':')[0].split('.')[-1]`` (run.py:36,105) still yields 'native'."""
import numpy as np

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
        # the history plot_state() draws (reinmav_env.py:80-84,119-124): one row per step, host side, N = 1 view only
        self.cum_state = self.stateToQd(self.state)
        self.cum_desired_state = np.zeros(11)
        self.cum_t = [0.0]

    @property
    def t(self):
        return self._batch.sim_time

    @t.setter
    def t(self, value):
        self._batch.sim_time = value

    def step(self, action=None):
        t_next = self.t + self.dt
        self._batch.step_host(None, self._rew, self._done, self._obs, auto_reset=False)
        state = self._obs[0].astype(float)
        self.cum_desired_state = np.vstack([self.cum_desired_state, self.trj_gen(t_next)])
        self.cum_state = np.vstack([self.cum_state, self.stateToQd(state)])
        self.cum_t.append(self.t)
        return state, float(self._rew[0]), bool(self._done[0]), {}

    @staticmethod
    def trj_gen(t):
        """The quintic 0 -> 1 in 4 s every axis and the yaw track (reinmav_env.py:128-136):
        [x, y, z, dx, dy, dz, ddx, ddy, ddz, yaw, dyaw]."""
        T = 4.0
        tau = min(max(float(t), 0.0), T) / T
        pos = 10.0 * tau ** 3 - 15.0 * tau ** 4 + 6.0 * tau ** 5
        vel = (30 / T) * tau ** 2 - (60 / T) * tau ** 3 + (30 / T) * tau ** 4
        acc = (60 / T ** 2) * tau - (180 / T ** 2) * tau ** 2 + (120 / T ** 2) * tau ** 3
        return [pos] * 3 + [vel] * 3 + [acc] * 3 + [pos, vel]

    @staticmethod
    def stateToQd(s):
        """13-state -> [pos(3), vel(3), roll, pitch, yaw, body rates(3)], ZXY Euler angles of the attitude
        (reinmav_env.py:292-304 with quat2mat :267-290 and RotToRPY :341-346)."""
        s = np.asarray(s, dtype=float)
        w, x, y, z = s[6:10]
        nq = w * w + x * x + y * y + z * z
        if nq > np.finfo(float).eps:
            k = 2.0 / nq
            r02, r10, r11 = k * (x * z + w * y), k * (x * y + w * z), 1.0 - k * (x * x + z * z)
            r12, r22 = k * (y * z - w * x), 1.0 - k * (x * x + y * y)
        else:
            r02 = r10 = r12 = 0.0
            r11 = r22 = 1.0
        phi = np.arcsin(r12)
        c = np.cos(phi)
        qd = np.empty(12)
        qd[:6] = s[:6]
        qd[6:9] = phi, np.arctan2(-r02 / c, r22 / c), np.arctan2(-r10 / c, r11 / c)
        qd[9:] = s[10:13]
        return qd

    def plot_state(self):
        """The reference's four figures (reinmav_env.py:138-186: x, dx and yaw against their references, saved as PDFs).
        Needs matplotlib, which the product does not depend on."""
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:
            raise ImportError("plot_state() needs matplotlib; the history is in cum_t / cum_state / cum_desired_state") from e
        figures = (("position_plot.pdf", 0, 0, "m", ["position x", "desired x"]),
                   ("velocity_plot.pdf", 3, 3, "m/s", ["velocity x", "desired vel x"]),
                   ("yaw_plot.pdf", 8, 9, "rad", ["yaw x", "desired yaw"]),
                   ("3dposition_plot.pdf", 0, 0, "m", ["position x", "desired x"]))
        for k, (name, col, ref_col, unit, legend) in enumerate(figures):
            fig = plt.figure(k + 1)
            plt.plot(self.cum_t, self.cum_state[:, col], "b", self.cum_t, self.cum_desired_state[:, ref_col], "r-.")
            plt.title("title"); plt.xlabel("Time(s)"); plt.ylabel(unit); plt.legend(legend); plt.grid(True)
            fig.savefig(name, format="pdf")
        plt.show()

    def control(self):
        raise TypeError("reinmav-v0 has no action input: its controller runs inside step()")
