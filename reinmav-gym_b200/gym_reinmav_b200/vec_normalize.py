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
        self._h_views = None
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
    def _host(self):
        """torch views of the env's page-locked host buffers + a device action buffer: the host protocol moves its arrays
        with asynchronous copies and one synchronisation per call, as QuadVecEnv.step does through rmv_step_host."""
        if self._h_views is None:
            import torch
            v = self.venv
            self._h_views = dict(act=torch.from_numpy(v._act.array), obs=torch.from_numpy(v._obs.array),
                                 rew=torch.from_numpy(v._rew.array), done=torch.from_numpy(v._done.array),
                                 act_dev=torch.empty(v._act.array.shape, dtype=torch.float32, device="cuda:%d" % self.device))
        return self._h_views

    def _sync(self):
        import torch
        torch.cuda.current_stream(self.device).synchronize()

    def reset(self):
        h = self._host()
        h["obs"].copy_(self.reset_tensor(), non_blocking=True)
        self._sync()
        return self.venv._out(self.venv._obs.array)

    def step_async(self, actions):
        h = self._host()
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != self.venv._act.array.shape:
            raise ValueError("actions must have shape %s, got %s" % (self.venv._act.array.shape, a.shape))
        if actions is not self.venv._act.array:
            self.venv._act.array[...] = a
        h["act_dev"].copy_(h["act"], non_blocking=True)
        self._act = h["act_dev"]

    def step_wait(self):
        if self._act is None:
            raise RuntimeError("step_wait() without step_async()")
        obs, rew, done = self.step_tensor(self._act)
        self._act = None
        h, v = self._host(), self.venv
        h["obs"].copy_(obs, non_blocking=True); h["rew"].copy_(rew, non_blocking=True); h["done"].copy_(done, non_blocking=True)
        self._sync()
        done_h = v._done.array.astype(bool)
        infos = v._finished(done_h) if v.monitor else SparseInfos(self.num_envs)
        return v._out(v._obs.array), v._out(v._rew.array), done_h, infos

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
        self._h_views = None                           # views of buffers the env frees
        lib().rmv_norm_destroy(self._h)
        self._h = ctypes.c_void_p()
        self.venv.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
