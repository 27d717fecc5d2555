"""CPU restatement of baselines' ``VecNormalize`` / ``RunningMeanStd``.

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

Third-party algorithm: the wrapper the reference's launcher puts around a vec-env
(/root/reference/gym_reinmav/run.py:91-92, ``env = VecNormalize(env)``) lives in openai/baselines, pinned by the
reference at commit 90d66776 (/root/reference/README.md:20,27), not vendored and not installed here.  Restated from its
published source, ``baselines/common/vec_env/vec_normalize.py`` and ``baselines/common/running_mean_std.py``:

    class RunningMeanStd:  mean = 0, var = 1, count = 1e-4 (float64); update(x): batch mean / np.var / count merged by
                           update_mean_var_count_from_moments (Chan et al.)
    class VecNormalize:    step_wait: ret = ret*gamma + rews; obs = _obfilt(obs); ret_rms.update(ret);
                           rews = clip(rews / sqrt(ret_rms.var + epsilon), -cliprew, cliprew); ret[news] = 0
                           _obfilt: ob_rms.update(obs); clip((obs - ob_rms.mean) / sqrt(ob_rms.var + epsilon), +-clipob)
                           reset: ret = zeros; _obfilt(venv.reset())

Parity unpinned by the package itself (absent); pinned to its algebra by tests/test_vecnorm.py: merging batches equals
the statistics of their concatenation, to 1e-12.
"""
import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        x = np.asarray(x, dtype=np.float64)
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, m2 / tot_count, tot_count


class VecNormalizeOracle:
    def __init__(self, num_envs, obs_dim, ob=True, ret=True, clipob=10.0, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        self.ob_rms = RunningMeanStd(shape=(obs_dim,)) if ob else None
        self.ret_rms = RunningMeanStd(shape=()) if ret else None
        self.clipob, self.cliprew, self.gamma, self.epsilon = clipob, cliprew, gamma, epsilon
        self.ret = np.zeros(num_envs)
        self.training = True

    def _obfilt(self, obs):
        if self.ob_rms is None:
            return obs
        if self.training:
            self.ob_rms.update(obs)
        return np.clip((obs - self.ob_rms.mean) / np.sqrt(self.ob_rms.var + self.epsilon), -self.clipob, self.clipob)

    def step(self, obs, rews, news):
        self.ret = self.ret * self.gamma + rews
        obs = self._obfilt(obs)
        if self.ret_rms is not None:
            if self.training:
                self.ret_rms.update(self.ret)
            rews = np.clip(rews / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[news] = 0.0
        return obs, rews

    def reset(self, obs):
        self.ret = np.zeros_like(self.ret)
        return self._obfilt(obs)
