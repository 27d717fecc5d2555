"""On-device rollout buffers in baselines-PPO2 shape.

baselines' ``ppo2.Runner.run`` (behind ``learn(env=env, ...)`` at /root/reference/gym_reinmav/run.py:63-68)
fills ``mb_obs, mb_actions, mb_rewards, mb_dones`` of shape ``[nsteps, nenv, ...]`` by looping
``actions = model.step(obs); obs, rewards, dones, infos = env.step(actions)``.  Here the buffers live in
HBM and the step kernel writes straight into them: step t reads ``actions[t]``, writes ``rewards[t]``,
``dones[t]`` and the observation rows ``obs[t + 1]`` — no staging copy, no host hop.

``travelling``: obs == state in the reference (quadrotor3d.py:104), so the rollout buffer itself can carry the
state — step t READS its state from ``obs[t]`` and writes the new one to ``obs[t + 1]`` (``rmv_step_rows``); the
handle's tiled state is not streamed at all during the rollout and is re-synchronised from ``obs[T]`` once at the
end.  Per env-step that is 4(2S+A+1)+1 bytes instead of 4(3S+A+1)+1 (the state is written once, not twice), at the
price of a shared-memory transpose on the load side as well.  Measured at N = 2^20 (profiles/r1b_sweep_obs_rows.txt):
quadrotor3d 15.3 vs 19.5 us per step, quadrotor3d-slungload 29.5 vs 31.9, quadrotor2d-slungload 17.9 vs 18.1,
quadrotor2d 11.4 vs 11.0 — so the default (``None``) turns it on for the 3D kinds only.  Same results bit for bit
either way (tests/test_gpu_vec_env.py).
"""
import torch


class RolloutBuffer:
    def __init__(self, venv, n_steps, travelling=None):
        self.venv, self.T = venv, int(n_steps)
        if travelling is None:
            # quadrotor3d, quadrotor3d-slungload — and any env that keeps its state as rows anyway
            travelling = venv.batch.state_dim in (10, 16) or getattr(venv, "state_layout", "tiles") == "rows"
        self.travelling = bool(travelling) and venv.batch.action_dim > 0
        n, S, A = venv.num_envs, venv.batch.state_dim, venv.batch.action_dim
        dev = "cuda:%d" % venv.device
        self.obs = torch.empty((self.T + 1, n, S), dtype=torch.float32, device=dev)      # obs[t] is what the policy sees at step t
        self.actions = torch.empty((self.T, n, A), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((self.T, n), dtype=torch.float32, device=dev)
        self.dones = torch.empty((self.T, n), dtype=torch.uint8, device=dev)
        self._have_first = False
        self._carry = False                                          # obs[T] of the last collect() is not yet in obs[0]
        # per-step views and raw addresses, built once: the collect() loop then costs one policy call and one
        # C-ABI call per step (at N = 2^20 a quadrotor2d step kernel takes ~10 us; seven tensor-indexing
        # operations per step cost more than that on the host)
        self.obs_views = [self.obs[t] for t in range(self.T + 1)]
        self.action_views = [self.actions[t] for t in range(self.T)]
        self._obs_p = [v.data_ptr() for v in self.obs_views]
        self._act_p = [v.data_ptr() for v in self.action_views]
        self._rew_p = [self.rewards[t].data_ptr() for t in range(self.T)]
        self._done_p = [self.dones[t].data_ptr() for t in range(self.T)]

    def start(self):
        """reset() all envs; the first observation lands in ``obs[0]``."""
        self.venv.batch.reset(obs=self.obs[0])
        self._have_first, self._carry = True, False
        return self.obs[0]

    def collect(self, policy):
        """Run ``T`` steps.  ``policy(obs[N,S]) -> actions[N,A]`` (CUDA tensors); its output is copied into
        ``actions[t]`` unless it already is that slice.  Returns self; ``obs[T]`` is the bootstrap observation.

        A second ``collect()`` continues where the first one ended, as ppo2's Runner does by keeping ``self.obs``:
        the bootstrap observation is carried into ``obs[0]`` first (``roll()``), whether or not the caller did so."""
        if not self._have_first:
            self.start()
        if self._carry:
            self.roll()
        b = self.venv.batch
        stream = torch.cuda.current_stream(self.venv.device).cuda_stream
        obs_t, act_t, obs_p, act_p, rew_p, done_p = self.obs_views, self.action_views, self._obs_p, self._act_p, self._rew_p, self._done_p
        if self.travelling:
            for t in range(self.T):
                a = policy(obs_t[t])
                if a.data_ptr() != act_p[t]:
                    act_t[t].copy_(a)
                b.step_rows(obs_p[t], act_p[t], rew_p[t], done_p[t], obs_p[t + 1], auto_reset=True, stream=stream)
            b.set_state_device(obs_p[self.T], stream=stream)       # the handle's own state follows the rollout
            self._carry = True
            return self
        for t in range(self.T):
            a = policy(obs_t[t])
            if a.data_ptr() != act_p[t]:
                act_t[t].copy_(a)
            b.step(act_p[t], rew_p[t], done_p[t], obs=obs_p[t + 1], auto_reset=True, stream=stream)
        self._carry = True
        return self

    def roll(self):
        """Make the bootstrap observation the first one of the next rollout (PPO2 keeps ``self.obs``)."""
        if self._carry:
            self.obs[0].copy_(self.obs[self.T])
            self._carry = False
