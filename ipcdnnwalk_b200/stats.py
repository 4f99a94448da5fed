"""Observation-normaliser / episode statistics: the one collective of the SRB hot path.

Reference: `VecNormalize(norm_obs=True)` keeps a `RunningMeanStd` in the trainer's main process
(gym_cdm2/train_gym.py:98, saved as obs_rms :230-233).  Sharded over GPUs, every rank accumulates the
sufficient statistics (count, sum x, sum x^2) of its own environments on the device
(`srb_obs_stats` kernel) and one NCCL all-reduce (sum) per rollout merges them -- O(300) doubles.
"""
import torch


class ObsStats:
    def __init__(self, dim, device):
        self.dim = dim
        self.acc = torch.zeros(1 + 2 * dim, dtype=torch.float64, device=device)   # this rank, since last merge
        self.total = torch.zeros(1 + 2 * dim, dtype=torch.float64, device=device)  # merged over ranks and rollouts

    def accumulate(self, env):
        """Add the env's current observation batch (CUDA kernel, asynchronous)."""
        env.obs_stats(self.acc)

    def merge(self, group=None):
        """All-reduce this rollout's sums over ranks and fold them into the running total."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=group)
        self.total += self.acc
        self.acc.zero_()

    @property
    def count(self):
        return float(self.total[0])

    def mean_var(self):
        n = torch.clamp(self.total[0], min=1.0)
        mean = self.total[1:1 + self.dim] / n
        var = torch.clamp(self.total[1 + self.dim:] / n - mean * mean, min=0.0)
        return mean, var

    def normalize(self, obs, clip=10.0, eps=1e-8):
        mean, var = self.mean_var()
        return torch.clamp((obs - mean.to(obs.dtype)) / torch.sqrt(var.to(obs.dtype) + eps), -clip, clip)
