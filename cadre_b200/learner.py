"""Data-parallel learner: the B200 replacement of the worker/chief handshake
(ppo_agent/train.py:76-110, ppo_agent/chief.py:8-27, ppo_agent/models.py:219-258).

Contract kept from the reference: every logical worker contributes exactly one gradient per update step;
the step applied is Adam on the SUM of the workers' gradients with a per-module clip at max_grad_norm; all
replicas hold identical parameters afterwards. How it is done here:
  * the W_local workers of a rank are processed by ONE batched cadre_ppo_update call (their gradients are
    summed on the device by construction);
  * across ranks the flat fp32 gradient (77.9 MB) is summed with one NCCL all-reduce over NVLink
    (`torch.distributed.all_reduce`), which replaces Shared_grad_buffers + the 1 Hz polling chief;
  * every rank then runs the same deterministic clip + Adam kernel, so no parameter broadcast
    (agent.py:239-243) is needed and replicas stay bit-identical.
"""
import os

import numpy as np
import torch

from . import ppo as _ppo
from ._lib import CadreError
from . import ppo_params
from .storage import RolloutStorage


class RolloutPool:
    """W x 2 RolloutStorage objects (steer, throttle per worker; train.py:44-48) whose tensors are views of
    batched device tensors, so that one GAE launch covers all 2W sequences."""

    def __init__(self, workers, rollout_cfg, device):
        self.workers, self.device = workers, torch.device(device)
        cfg = dict(rollout_cfg)
        cfg.setdefault("hidden_size", cfg["feature_dims"])
        T, S, Fd, Hd = cfg["num_steps"], cfg["seq_length"], cfg["feature_dims"], cfg["hidden_size"]
        n = workers * 2
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=self.device)  # noqa: E731
        self.batched = dict(
            command=z(n, T + 1, 1, dtype=torch.int32), obs=z(n, T + 1, S, Fd), rewards=z(n, T + 1, 1),
            value_preds=z(n, T + 1, 1), returns=z(n, T + 1, 1), action_log_probs=z(n, T + 1, 1),
            action=z(n, T + 1, 1, dtype=torch.int64), masks=z(n, T + 1, 1), hn=z(n, T + 1, Hd), cn=z(n, T + 1, Hd),
            advantages=z(n, T, 1))
        self.next_value = z(n)
        self.storages = []
        for w in range(workers):
            pair = []
            for h in range(2):
                st = RolloutStorage(**cfg)
                for name, t in self.batched.items():
                    setattr(st, name, t[w * 2 + h])
                pair.append(st)
            self.storages.append(tuple(pair))
        self.num_steps, self.gamma, self.tau = T, cfg["gamma"], cfg["tau"]

    def compute_returns(self, next_values=None, normalize=True):
        """All 2W sequences in one launch (storage.py:68-76 + train.py:82-88). next_values: [W,2] or None to use
        self.next_value."""
        if next_values is not None:
            self.next_value.copy_(torch.as_tensor(next_values, dtype=torch.float32).reshape(-1))
        b, T = self.batched, self.num_steps
        n = self.workers * 2
        _ppo.gae(b["rewards"].view(n, T + 1), b["value_preds"].view(n, T + 1), b["masks"].view(n, T + 1),
                 self.next_value, b["returns"].view(n, T + 1), b["advantages"].view(n, T), self.gamma, self.tau,
                 normalize)


class Learner:
    def __init__(self, workers, mini_batch, ppo_state, device="cuda:0", clip=0.1, value_coeff=0.1, clip_coeff=1.0,
                 ent_coeff=0.01, lr=3e-4, max_grad_norm=250.0, process_group=None, seeds=None):
        self.device = torch.device(device)
        self.workers, self.mini_batch = workers, mini_batch
        self.value_coeff, self.clip_coeff, self.ent_coeff = value_coeff, clip_coeff, ent_coeff
        self.lr, self.max_grad_norm = lr, max_grad_norm
        self.engine = _ppo.PpoEngine(workers, mini_batch, clip, value_coeff, clip_coeff, ent_coeff, self.device)
        self.params = ppo_params.pack_state(ppo_state, self.device)
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        # Gradient exchange: "switch" = the library's in-switch reduction on a symmetric gradient buffer
        # (collective.SwitchAllReduce, csrc/allreduce.cu), "nccl" = torch.distributed.all_reduce. "auto" (default) takes
        # the switch from 4 ranks on when the fabric offers a multicast mapping: measured on 8 x B200 the 77.9 MB sum
        # takes 0.20 ms in the switch against 0.32 ms (NCCL default) / 0.27 ms (NCCL ring); between 2 GPUs NCCL's
        # ring is the faster one in situ.
        self.exchange = os.environ.get("CADRE_ALLREDUCE", "auto") if self.world > 1 else "none"
        if self.exchange not in ("auto", "switch", "nccl", "none"):
            raise CadreError(f"CADRE_ALLREDUCE={self.exchange!r}: expected 'auto', 'switch' or 'nccl'")
        self._switch = None
        if self.exchange == "switch" or (self.exchange == "auto" and self.world >= 4):
            from .collective import SwitchAllReduce
            try:
                self._switch = SwitchAllReduce(self.params.numel(), self.device, process_group)
            except Exception as e:     # no symmetric memory between these devices (not one NVLink domain, two ranks per GPU)
                if self.exchange == "switch":
                    raise
                import warnings
                warnings.warn(f"cadre_b200: in-switch all-reduce unavailable ({e}); using NCCL")
            if self._switch is not None and self.exchange == "auto" and not self._switch.multicast:
                self._switch = None    # peer loads / stores only pay off between two GPUs
        if self._switch is not None:
            self.exchange = "switch"
            self.grads = self._switch.buffer
        else:
            self.exchange = "nccl" if self.world > 1 else "none"
            self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.losses = torch.zeros(workers, 2, 3, device=self.device)       # last update step
        self.loss_sum = torch.zeros(workers, 2, 3, device=self.device)     # summed over the update steps of learn()
        self.loss_steps = 0
        self.step_count = 0
        self.pg = process_group
        self.overlap_allreduce = os.environ.get("CADRE_NO_ALLREDUCE_OVERLAP", "0") != "1"
        self._comm_stream = None
        # one CPU RNG stream per logical worker (each reference worker is its own process with its own global RNG)
        self._rng_states = []
        if seeds is not None:
            keep = torch.get_rng_state()
            for s in seeds:
                torch.manual_seed(int(s))
                self._rng_states.append(torch.get_rng_state())
            torch.set_rng_state(keep)

    # ---------------------------------------------------------------- index sampling (bit-exact, host side)
    def sample_epoch_indices(self, storages):
        """One PPO epoch of minibatch indices for every worker: [n_minibatches][W][2][mb] (train.py:94-96: the
        steer generator draws its permutation first, then the throttle generator)."""
        per_worker = []
        keep = torch.get_rng_state() if self._rng_states else None
        for w, (st_s, st_t) in enumerate(storages):
            if self._rng_states:
                torch.set_rng_state(self._rng_states[w])
            s_chunks = st_s.sample_indices()
            t_chunks = st_t.sample_indices()
            if self._rng_states:
                self._rng_states[w] = torch.get_rng_state()
            per_worker.append((s_chunks, t_chunks))
        if keep is not None:
            torch.set_rng_state(keep)
        n_mb = min(min(len(s), len(t)) for s, t in per_worker)     # zip() of the two generators (train.py:94)
        out = np.empty((n_mb, len(storages), 2, self.mini_batch), dtype=np.int32)
        for w, (s_chunks, t_chunks) in enumerate(per_worker):
            for k in range(n_mb):
                for h, chunk in enumerate((s_chunks[k], t_chunks[k])):
                    if len(chunk) != self.mini_batch:
                        # BatchSampler(drop_last=False) yields a short last chunk when num_steps is not a multiple of
                        # num_steps // mini_batch_num (storage.py:93-97); the batched engine is sized for ONE
                        # minibatch length. CadreAgent.update_policy handles such tails (one engine per length).
                        raise CadreError(
                            f"ragged minibatch: chunk {k} of worker {w} has {len(chunk)} rows, the learner was built "
                            f"for mini_batch={self.mini_batch}; choose num_steps divisible by num_steps // "
                            "mini_batch_num or drive CadreAgent.update_policy per worker")
                    out[k, w, h] = chunk
        return out

    # ---------------------------------------------------------------- one synchronous update step
    def _setup_pipeline(self):
        """Streams, events and flat-buffer ranges of the data-parallel gradient pipeline (built on first use).
        Range k < groups - 1 = LSTM tensors of expert group k; the last range = LSTM tensors of the last group PLUS the
        actor-critic tensors of all experts, which follow them directly in the flat buffer (one collective less, and
        no collective is in flight next to the persistent BPTT kernel, which needs every SM it was launched on)."""
        # NCCL: two expert groups (range 0 is on the wire while group 1's weight-gradient GEMMs run); the in-switch
        # exchange measured no gain from splitting (8 x B200: PPO phase 7.86 ms with two ranges, 7.80 ms with one)
        default_groups = "1" if self._switch is not None else "2"
        groups = int(os.environ.get("CADRE_GRAD_GROUPS", default_groups)) if self.overlap_allreduce else 1
        self.engine.set_grad_groups(groups)
        self._comm_stream = torch.cuda.Stream(device=self.device)
        lstm = [self.engine.grad_range(g) for g in range(groups)]
        mlp = self.engine.grad_range(-1)
        assert lstm[-1][0] + lstm[-1][1] == mlp[0] and lstm[-1][3] == mlp[2]      # contiguous ranges and module ids
        self._ranges = lstm[:-1] + [(lstm[-1][0], lstm[-1][1] + mlp[1], lstm[-1][2], mlp[3])]
        self._waits = [[g] for g in range(groups - 1)] + [[groups - 1, -1]]
        self._reduced = [torch.cuda.Event() for _ in self._ranges]
        assert sum(r[1] for r in self._ranges) == self.grads.numel()
        assert all(r[0] % 4 == 0 and r[1] % 4 == 0 for r in self._ranges)        # 16-byte vectors in the exchange

    def _exchange_and_step(self, step):
        """all-reduce(sum) + per-module clip + Adam on the gradient the last engine.update left in self.grads.
        `step`: 1-based Adam step, or 0 to use the device-side counter (staged sequences of update steps).

        With more than one rank the gradient exchange is pipelined against the end of the backward pass: the LSTM
        weight gradients (72 of the 78 MB) are produced per group of experts, each group one contiguous range of the
        flat buffer; a communication stream all-reduces range k as soon as its event fires while the GEMMs of group
        k+1 still run, and the main stream applies clip + Adam to the modules of range k - a per-module operation,
        chief.py:16-21 - while range k+1 is on the wire. Sums, clip coefficients and Adam arithmetic are those of the
        single all-reduce + single step."""
        if self.world > 1:
            main = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(self._comm_stream):
                for k, (off, cnt, _, _) in enumerate(self._ranges):
                    for g in self._waits[k]:
                        self.engine.wait_grads(g, self._comm_stream)
                    if self._switch is not None:
                        self._switch.sum_(off, cnt, stream=self._comm_stream)
                    else:
                        torch.distributed.all_reduce(self.grads[off:off + cnt], op=torch.distributed.ReduceOp.SUM,
                                                     group=self.pg)
                    self._reduced[k].record(self._comm_stream)
            for k, (_, _, m0, m1) in enumerate(self._ranges):
                main.wait_event(self._reduced[k])
                self.engine.adam_step_modules(self.params, self.grads, self.exp_avg, self.exp_avg_sq, step, m0, m1,
                                              self.max_grad_norm, self.lr)
        else:
            self.engine.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, step, self.max_grad_norm,
                                  self.lr)
        self.loss_sum += self.losses
        self.loss_steps += 1

    def update_step(self, storages, indices, async_losses=True):
        """update_policy for all local workers -> all-reduce(sum) -> per-module clip + Adam (one step, eagerly).
        storages[w] = (steer, throttle) RolloutStorage with `.advantages`; indices int32 [W,2,mb]."""
        advs = [(s.advantages, t.advantages) for s, t in storages]
        if self.world > 1 and self._comm_stream is None:
            self._setup_pipeline()
        self.engine.update(storages, advs, indices, self.params, self.grads, self.losses)
        self.step_count += 1
        self._exchange_and_step(self.step_count)
        return self.losses if async_losses else self.scaled_losses()

    def scaled_losses(self, mean=False):
        """[W,3] = (value_loss*coeff, action_loss*coeff, entropy*coeff) per worker, like update_policy's return:
        of the LAST update step, or with `mean=True` averaged over the update steps since the last `learn()` began
        (what train.py:98-100, 112-116 logs)."""
        L = (self.loss_sum / max(1, self.loss_steps) if mean else self.losses).sum(1).cpu()
        return L * torch.tensor([self.value_coeff, self.clip_coeff, self.ent_coeff])

    # ---------------------------------------------------------------- a whole learn() phase
    def learn(self, pool_or_storages, ppo_epoch=4):
        """ppo_epoch x minibatches of update steps over already computed returns / advantages (train.py:93-110).
        All minibatch indices of the phase are drawn first (same per-worker RNG streams and order as one generator
        pair per epoch, train.py:94) and uploaded once together with the storage references; every update step then
        consumes the next slice of that device-side table (and the device-side Adam step), so no host data moves
        between the steps of a phase."""
        storages = pool_or_storages.storages if hasattr(pool_or_storages, "storages") else pool_or_storages
        self.loss_sum.zero_()
        self.loss_steps = 0
        if self.world > 1 and self._comm_stream is None:
            self._setup_pipeline()
        idx_all = np.concatenate([self.sample_epoch_indices(storages) for _ in range(ppo_epoch)], 0)
        advs = [(s.advantages, t.advantages) for s, t in storages]
        n = 0
        while n < idx_all.shape[0]:                       # (the staging table holds 64 steps)
            m = min(64, idx_all.shape[0] - n)
            self.engine.stage(storages, advs, idx_all[n:n + m], self.step_count + 1)
            for _ in range(m):
                self.engine.update_staged(self.params, self.grads, self.losses)
                self._exchange_and_step(0)
                self.step_count += 1
            n += m
        self.loss_steps = idx_all.shape[0]
        return idx_all.shape[0]

    def check(self):
        """Synchronise and raise CadreError if a bounded wait inside a kernel (LSTM hand-off, all-reduce barrier) ever
        timed out."""
        self.engine.check()
        if self._switch is not None:
            self._switch.check()

    def state(self):
        return ppo_params.unpack_state(self.params)
