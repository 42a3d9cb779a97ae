"""RolloutStorage drop-in (reference: ppo_agent/storage.py:5-120).

Same constructor, tensor attributes, `insert` / `after_update` / `get_last` semantics (including the cursor that
wraps modulo num_steps+1 because the reference never calls `after_update`) and the same
minibatch index stream: `feed_forward_generator` draws `BatchSampler(SubsetRandomSampler(range(T)), T // n)`
from the global CPU generator, so indices are bit-exact with the reference for the same torch seed.

Differences are on the device side only: `compute_returns` runs the fused GAE scan kernel (cadre_gae) and also
produces the normalised advantages of train.py:82-88 in the same pass (`self.advantages`), and the generator
yields `MiniBatch` objects that carry (storage, indices) for the fused gather inside the PPO update while still
unpacking to the reference's 9-tuple of tensors.
"""
import torch
from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler

from . import ppo as _ppo


class MiniBatch:
    """One minibatch of a RolloutStorage. Iterating / indexing gives the reference's 9-tuple
    (obs, action, value_preds, returns, masks, old_log_probs, advantages, [hn, cn], command), materialised
    lazily with torch indexing; the CUDA update path uses `.storage`, `.indices`, `.advantages` directly."""

    def __init__(self, storage, indices, advantages):
        self.storage, self.indices, self.advantages = storage, list(indices), advantages
        self._tuple = None

    def as_tuple(self):
        if self._tuple is None:
            s, idx = self.storage, self.indices
            obs = s.obs[idx].permute(1, 0, 2)
            obs = obs.reshape(-1, obs.size(-1))
            self._tuple = (obs, s.action[idx], s.value_preds[idx], s.returns[idx], s.masks[idx],
                           s.action_log_probs[idx], self.advantages[idx], [s.hn[idx], s.cn[idx]], s.command[idx])
        return self._tuple

    def __iter__(self):
        return iter(self.as_tuple())

    def __getitem__(self, i):
        return self.as_tuple()[i]

    def __len__(self):
        return 9


class RolloutStorage(object):
    def __init__(self, num_steps, mini_batch_num, feature_dims, seq_length, hidden_size, use_gae, gamma, tau):
        self.mini_batch_num = mini_batch_num
        self.command = torch.zeros((num_steps + 1, 1), dtype=torch.int)
        self.obs = torch.zeros(num_steps + 1, seq_length, feature_dims)
        self.z_dims = feature_dims
        self.rewards = torch.zeros(num_steps + 1, 1)
        self.value_preds = torch.zeros(num_steps + 1, 1)
        self.returns = torch.zeros(num_steps + 1, 1)
        self.action_log_probs = torch.zeros(num_steps + 1, 1)
        self.action = torch.zeros((num_steps + 1, 1), dtype=torch.long)
        self.seq_length = seq_length
        self.hn = torch.zeros(num_steps + 1, hidden_size)
        self.cn = torch.zeros(num_steps + 1, hidden_size)
        self.hid_size = hidden_size
        self.masks = torch.zeros(num_steps + 1, 1)
        self.advantages = torch.zeros(num_steps, 1)   # filled by compute_returns (train.py:82-88)
        self.num_steps = num_steps
        self.use_gae = use_gae
        self.gamma = gamma
        self.tau = tau
        self.step = 0

    _TENSORS = ("command", "obs", "rewards", "value_preds", "returns", "action_log_probs", "action", "masks", "hn",
                "cn", "advantages")

    def to(self, device):
        for name in self._TENSORS:
            setattr(self, name, getattr(self, name).to(device))

    def insert(self, obs, action, action_log_probs, value_preds, rewards, masks, hidden_state, command):
        # storage.py:45-58
        self.action[self.step].copy_(action.squeeze())
        self.action_log_probs[self.step].copy_(action_log_probs.squeeze())
        self.value_preds[self.step].copy_(value_preds.squeeze())
        self.rewards[self.step].copy_(rewards.squeeze())
        self.obs[self.step].copy_(obs.squeeze())
        if hidden_state is not None and self.step < self.num_steps:
            hn, cn = hidden_state
            self.hn[self.step + 1].copy_(hn.clone().squeeze())
            self.cn[self.step + 1].copy_(cn.clone().squeeze())
        self.masks[self.step].copy_(masks.squeeze())
        self.command[self.step] = command
        self.step = (self.step + 1) % (self.num_steps + 1)

    def after_update(self, hidden_state):
        self.step = 0
        if hidden_state is not None:
            hn, cn = hidden_state
            self.hn[0].copy_(hn.squeeze())
            self.cn[0].copy_(cn.squeeze())

    def compute_returns(self, next_value, normalize=True):
        """storage.py:68-76 (GAE branch) on the device; also fills `self.advantages` (train.py:82-88)."""
        if not self.use_gae:
            raise NotImplementedError("only the use_gae=True branch (agent_config.py:22) is on the hot path")
        if not self.rewards.is_cuda:
            raise _ppo._lib.CadreError("RolloutStorage.compute_returns needs the storage on a CUDA device "
                                       "(call .to(device)); there is no CPU fallback")
        nv = torch.as_tensor(next_value, dtype=torch.float32).reshape(1).to(self.rewards.device)
        T = self.num_steps
        _ppo.gae(self.rewards.view(1, T + 1), self.value_preds.view(1, T + 1), self.masks.view(1, T + 1), nv,
                 self.returns.view(1, T + 1), self.advantages.view(1, T), self.gamma, self.tau, normalize)

    def get_last(self):
        return self.obs[-1], self.command[-1].item()

    def sample_indices(self):
        """The index chunks feed_forward_generator will use (one torch.randperm from the global CPU RNG)."""
        mini_batch_size = self.num_steps // self.mini_batch_num
        sampler = BatchSampler(SubsetRandomSampler(range(0, self.num_steps)), mini_batch_size, drop_last=False)
        return [list(ix) for ix in sampler]

    def feed_forward_generator(self, advantages=None):
        # storage.py:93-120; lazily sampled like the reference (the permutation is drawn at the first next())
        adv = self.advantages if advantages is None else advantages
        mini_batch_size = self.num_steps // self.mini_batch_num
        sampler = BatchSampler(SubsetRandomSampler(range(0, self.num_steps)), mini_batch_size, drop_last=False)
        for indices in sampler:
            yield MiniBatch(self, indices, adv)
