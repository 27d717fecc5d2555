"""VecNormalize — baselines' running observation / return normalisation, on the device.

The reference's launcher wraps a vec-env in ``VecNormalize`` when ``env_type == 'mujoco'``
(/root/reference/gym_reinmav/run.py:91-92); for the native envs it is an opt-in with the same constructor
(``VecNormalize(venv, ob=True, ret=True, clipob=10., cliprew=10., gamma=0.99, epsilon=1e-8)``) and the same protocol.
The statistics (float64) and the discounted returns live in HBM; one step costs two small kernels after the env step
(csrc/rmv_norm.cuh) and no host round trip.

* ``step_tensor(actions)`` / ``reset_tensor()`` — device in, device out (normalised obs / rewards alias internal buffers).
* ``step(actions)`` / ``reset()`` — NumPy in, NumPy out (the VecEnv protocol): the env steps on the device, the normalised
  results are copied to the host.
"""
import ctypes

import numpy as np

from ._capi import check, lib
from .batch import _ptr, _stream
from .vec_env import SparseInfos


class VecNormalize:
    def __init__(self, venv, ob=True, ret=True, clipob=10.0, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        import torch
        self.venv = venv
        self.num_envs, self.observation_space, self.action_space = venv.num_envs, venv.observation_space, venv.action_space
        self.metadata, self.device = venv.metadata, venv.device
        self.training = True                         # False freezes the statistics (evaluation)
        self._h = ctypes.c_void_p()
        S = venv.batch.state_dim
        check(lib().rmv_norm_create(ctypes.byref(self._h), self.num_envs, S, self.device, 1 if ob else 0, 1 if ret else 0,
                                    float(clipob), float(cliprew), float(gamma), float(epsilon)))
        dev = "cuda:%d" % self.device
        self._obs = torch.empty((self.num_envs, S), dtype=torch.float32, device=dev)
        self._rew = torch.empty((self.num_envs,), dtype=torch.float32, device=dev)
        self._obs_ptr, self._rew_ptr = self._obs.data_ptr(), self._rew.data_ptr()
        self._act = None
        self.closed = False

    # ------------------------------------------------------------------ device path
    def reset_tensor(self):
        raw = self.venv.reset_tensor()
        check(lib().rmv_norm_reset(self._h, _ptr(raw, "float32", self.device, "obs"), _ptr(self._obs), 1 if self.training else 0,
                                   _stream(None, self.device)))
        return self._obs

    def step_tensor(self, actions):
        """(normalised obs [N,S], normalised reward [N], done [N] uint8) — tensors aliasing internal buffers."""
        raw, rew, done = self.venv.step_tensor(actions)
        t = self.venv._tensors()                       # the env's own buffers: addresses validated when they were made
        check(lib().rmv_norm_step(self._h, t["obs_ptr"], t["rew_ptr"], t["done_ptr"], self._obs_ptr, self._rew_ptr,
                                  1 if self.training else 0, _stream(None, self.device)))
        return self._obs, self._rew, done

    # ------------------------------------------------------------------ VecEnv protocol (host)
    def reset(self):
        return self.reset_tensor().cpu().numpy()

    def step_async(self, actions):
        import torch
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, self.venv.batch.action_dim):
            raise ValueError("actions must have shape %s, got %s" % ((self.num_envs, self.venv.batch.action_dim), a.shape))
        self._act = torch.as_tensor(a, device="cuda:%d" % self.device)

    def step_wait(self):
        if self._act is None:
            raise RuntimeError("step_wait() without step_async()")
        obs, rew, done = self.step_tensor(self._act)
        self._act = None
        done_h = done.cpu().numpy().astype(bool)
        infos = self.venv._finished(done_h) if self.venv.monitor else SparseInfos(self.num_envs)
        return obs.cpu().numpy(), rew.cpu().numpy(), done_h, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ------------------------------------------------------------------ statistics (checkpoint / logging)
    def statistics(self):
        S = self.venv.batch.state_dim
        mean, var = (ctypes.c_double * S)(), (ctypes.c_double * S)()
        oc, rm, rv, rc = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(lib().rmv_norm_get(self._h, mean, var, ctypes.byref(oc), ctypes.byref(rm), ctypes.byref(rv), ctypes.byref(rc),
                                 _stream(None, self.device)))
        return {"ob_mean": np.array(mean), "ob_var": np.array(var), "ob_count": oc.value,
                "ret_mean": rm.value, "ret_var": rv.value, "ret_count": rc.value}

    def load_statistics(self, st):
        S = self.venv.batch.state_dim
        mean = (ctypes.c_double * S)(*[float(x) for x in st["ob_mean"]])
        var = (ctypes.c_double * S)(*[float(x) for x in st["ob_var"]])
        check(lib().rmv_norm_set(self._h, mean, var, float(st["ob_count"]), float(st["ret_mean"]), float(st["ret_var"]),
                                 float(st["ret_count"]), _stream(None, self.device)))

    def render(self, mode="human"):
        return self.venv.render(mode)

    def episode_stats(self):
        return self.venv.episode_stats()

    def close(self):
        if self.closed:
            return
        self.closed = True
        lib().rmv_norm_destroy(self._h)
        self._h = ctypes.c_void_p()
        self.venv.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
