"""On-device rollout buffers in baselines-PPO2 shape.

baselines' ``ppo2.Runner.run`` (behind ``learn(env=env, ...)`` at /root/reference/gym_reinmav/run.py:63-68)
fills ``mb_obs, mb_actions, mb_rewards, mb_dones`` of shape ``[nsteps, nenv, ...]`` by looping
``actions = model.step(obs); obs, rewards, dones, infos = env.step(actions)``.  Here the buffers live in
HBM and the step kernel writes straight into them: step t reads ``actions[t]``, writes ``rewards[t]``,
``dones[t]`` and the observation rows ``obs[t + 1]`` — no staging copy, no host hop.
"""
import torch


class RolloutBuffer:
    def __init__(self, venv, n_steps):
        self.venv, self.T = venv, int(n_steps)
        n, S, A = venv.num_envs, venv.batch.state_dim, venv.batch.action_dim
        dev = "cuda:%d" % venv.device
        self.obs = torch.empty((self.T + 1, n, S), dtype=torch.float32, device=dev)      # obs[t] is what the policy sees at step t
        self.actions = torch.empty((self.T, n, A), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((self.T, n), dtype=torch.float32, device=dev)
        self.dones = torch.empty((self.T, n), dtype=torch.uint8, device=dev)
        self._have_first = False

    def start(self):
        """reset() all envs; the first observation lands in ``obs[0]``."""
        self.venv.batch.reset(obs=self.obs[0])
        self._have_first = True
        return self.obs[0]

    def collect(self, policy):
        """Run ``T`` steps.  ``policy(obs[N,S]) -> actions[N,A]`` (CUDA tensors); its output is copied into
        ``actions[t]`` unless it already is that slice.  Returns self; ``obs[T]`` is the bootstrap observation."""
        if not self._have_first:
            self.start()
        b = self.venv.batch
        for t in range(self.T):
            a = policy(self.obs[t])
            if a.data_ptr() != self.actions[t].data_ptr():
                self.actions[t].copy_(a)
            b.step(self.actions[t], self.rewards[t], self.dones[t], obs=self.obs[t + 1], auto_reset=True)
        return self

    def roll(self):
        """Make the bootstrap observation the first one of the next rollout (PPO2 keeps ``self.obs``)."""
        self.obs[0].copy_(self.obs[self.T])
