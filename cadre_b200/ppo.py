"""ctypes front of the PPO update plan in libcadre_sm100.so (cadre_ppo_* in include/cadre_b200.h).

`PpoEngine` is the device-side replacement of `CadreAgent.update_policy` + `Shared_grad_buffers.add_gradient`
+ the `chief` step (ppo_agent/agent.py:166-237, models.py:231-239, chief.py:13-21) for W logical workers whose
rollouts live on this GPU. Host logic here is only pointer / index marshalling.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ppo_params


class PpoConfigC(ctypes.Structure):
    _fields_ = [("workers", ctypes.c_int32), ("mini_batch", ctypes.c_int32), ("clip", ctypes.c_float),
                ("value_coeff", ctypes.c_float), ("clip_coeff", ctypes.c_float), ("ent_coeff", ctypes.c_float)]


class StorageRefC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("obs", "action", "value_preds", "returns", "action_log_probs", "adv",
                                               "hn", "cn", "command")]


ROW_OUT = 36


def _check_tensor(t, dtype, name):
    if t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
        raise _lib.CadreError(f"storage tensor `{name}` must be a contiguous CUDA {dtype} tensor "
                              f"(got {t.dtype}, contiguous={t.is_contiguous()}, device={t.device})")


def storage_ref(st, adv):
    """cadre_storage_ref for one RolloutStorage-like object (attributes of ppo_agent/storage.py:8-26)."""
    r = StorageRefC()
    for name, dtype in (("obs", torch.float32), ("action", torch.int64), ("value_preds", torch.float32),
                        ("returns", torch.float32), ("action_log_probs", torch.float32), ("hn", torch.float32),
                        ("cn", torch.float32), ("command", torch.int32)):
        t = getattr(st, name)
        _check_tensor(t, dtype, name)
        setattr(r, name, t.data_ptr())
    _check_tensor(adv, torch.float32, "advantages")
    r.adv = adv.data_ptr()
    return r


class PpoEngine:
    def __init__(self, workers, mini_batch, clip=0.1, value_coeff=0.1, clip_coeff=1.0, ent_coeff=0.01,
                 device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CadreError("PpoEngine needs a CUDA device: there is no CPU fallback")
        self.workers, self.mini_batch = int(workers), int(mini_batch)
        self._lib = _lib.lib()
        self._lib.cadre_ppo_param_count.restype = ctypes.c_int64
        assert self._lib.cadre_ppo_param_count() == ppo_params.TOTAL, "ppo_layout.h / ppo_params.py mismatch"
        cfg = PpoConfigC(self.workers, self.mini_batch, clip, value_coeff, clip_coeff, ent_coeff)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_create(ctypes.byref(h), ctypes.byref(cfg)))
        self._h = h
        self._refs = (StorageRefC * (2 * self.workers))()
        self._keep = None
        self.grad_groups = 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.cadre_ppo_destroy(h)
            self._h = None

    def _marshal(self, storages, advantages, indices):
        assert len(storages) == self.workers and len(advantages) == self.workers
        for w in range(self.workers):
            for h in range(2):
                self._refs[w * 2 + h] = storage_ref(storages[w][h], advantages[w][h])
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32))
        assert idx.shape == (self.workers, 2, self.mini_batch), idx.shape
        self._keep = (storages, advantages, idx)  # keep alive until the stream has consumed them
        return idx

    def update(self, storages, advantages, indices, params, grads, losses=None):
        """storages[w] = (steer_storage, throttle_storage); advantages[w] = (adv_steer [T,1], adv_throttle);
        indices int [W,2,mb]. Overwrites `grads` with the sum over workers of d(total_loss)/d(params) and
        returns the device tensor losses [W,2,3] = un-scaled (value, action, entropy) means per worker/head."""
        idx = self._marshal(storages, advantages, indices)
        if losses is None:
            losses = torch.empty(self.workers, 2, 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_update(self._h, self._refs, idx.ctypes.data_as(ctypes.c_void_p),
                                                  _lib.ptr(params), _lib.ptr(grads), _lib.ptr(losses),
                                                  _lib.stream_ptr(self.device)))
        return losses

    def stage(self, storages, advantages, indices_all, first_adam_step):
        """Upload the storage references and the indices of a SEQUENCE of update steps, int [n_steps, W, 2, mb]; each
        following `update_staged` consumes the next slice and advances the device-side Adam step (the first one is
        `first_adam_step`, 1-based), which `adam_step*` use when called with step=0."""
        assert len(storages) == self.workers and len(advantages) == self.workers
        for w in range(self.workers):
            for h in range(2):
                self._refs[w * 2 + h] = storage_ref(storages[w][h], advantages[w][h])
        idx = np.ascontiguousarray(np.asarray(indices_all, dtype=np.int32))
        assert idx.ndim == 4 and idx.shape[1:] == (self.workers, 2, self.mini_batch), idx.shape
        self._keep = (storages, advantages, idx)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_stage(self._h, self._refs, idx.ctypes.data_as(ctypes.c_void_p), idx.shape[0],
                                                 int(first_adam_step), _lib.stream_ptr(self.device)))
        return idx.shape[0]

    def update_staged(self, params, grads, losses):
        """`update` on the next staged index slice: no host data, capturable in a CUDA graph."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_update(self._h, None, None, _lib.ptr(params), _lib.ptr(grads),
                                                  _lib.ptr(losses), _lib.stream_ptr(self.device)))
        return losses

    def evaluate(self, storages, advantages, indices, params):
        """Forward only: returns [2, W*mb, 36] = value, log-prob(action), entropy, 33 normalised logits."""
        idx = self._marshal(storages, advantages, indices)
        out = torch.zeros(2, self.workers * self.mini_batch, ROW_OUT, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_evaluate(self._h, self._refs, idx.ctypes.data_as(ctypes.c_void_p),
                                                    _lib.ptr(params), _lib.ptr(out), _lib.stream_ptr(self.device)))
        return out

    def adam_step(self, params, grads, exp_avg, exp_avg_sq, step, max_grad_norm=250.0, lr=3e-4, betas=(0.9, 0.999),
                  eps=1e-8):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_adam_step(
                self._h, _lib.ptr(params), _lib.ptr(grads), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                ctypes.c_float(max_grad_norm), ctypes.c_float(lr), ctypes.c_float(betas[0]), ctypes.c_float(betas[1]),
                ctypes.c_float(eps), int(step), _lib.stream_ptr(self.device)))

    def module_norms(self):
        """{module name: gradient norm seen by the last adam_step} (16 entries, reference module names)."""
        buf = (ctypes.c_float * 16)()
        _lib.check(self._lib.cadre_ppo_module_norms(self._h, buf))
        out = {}
        for e in range(8):
            head, c = ppo_params.HEADS[e // 4], e % 4
            out[f"{head}_lstm_{c}"] = buf[e]
            out[f"{head}_ppo_{c}"] = buf[8 + e]
        return out

    # ---- data-parallel pipeline (see include/cadre_b200.h: cadre_ppo_set_grad_groups ...)
    def set_grad_groups(self, groups):
        _lib.check(self._lib.cadre_ppo_set_grad_groups(self._h, int(groups)))
        self.grad_groups = int(groups)

    def grad_range(self, group):
        """(offset, count, mod_begin, mod_end) of group `group`'s LSTM tensors in the flat buffers; group = -1: the
        actor-critic tensors of all experts."""
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        m0, m1 = ctypes.c_int(), ctypes.c_int()
        _lib.check(self._lib.cadre_ppo_grad_range(self._h, int(group), ctypes.byref(off), ctypes.byref(cnt),
                                                  ctypes.byref(m0), ctypes.byref(m1)))
        return off.value, cnt.value, m0.value, m1.value

    def wait_grads(self, group, stream):
        """`stream` (torch.cuda.Stream) waits until the last update() has finished that range of the gradient."""
        _lib.check(self._lib.cadre_ppo_wait_grads(self._h, int(group), ctypes.c_void_p(stream.cuda_stream)))

    def adam_step_modules(self, params, grads, exp_avg, exp_avg_sq, step, mod_begin, mod_end, max_grad_norm=250.0,
                          lr=3e-4, betas=(0.9, 0.999), eps=1e-8):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_adam_step_modules(
                self._h, _lib.ptr(params), _lib.ptr(grads), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                ctypes.c_float(max_grad_norm), ctypes.c_float(lr), ctypes.c_float(betas[0]), ctypes.c_float(betas[1]),
                ctypes.c_float(eps), int(step), int(mod_begin), int(mod_end), _lib.stream_ptr(self.device)))

    def check(self):
        """Synchronise and raise CadreError if a recurrence kernel of an earlier call timed out on a hand-off."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_ppo_check(self._h))

    @property
    def launches(self):
        return int(self._lib.cadre_ppo_launches(self._h))


def gae(rewards, values, masks, next_value, returns, adv, gamma=0.99, tau=0.95, normalize=True):
    """cadre_gae over E sequences: rewards/values/masks/returns fp32 [E,T+1], next_value [E], adv [E,T]."""
    E, T1 = rewards.shape
    for t in (rewards, values, masks, next_value, returns, adv):
        _check_tensor(t, torch.float32, "gae input")
    with torch.cuda.device(rewards.device):
        _lib.check(_lib.lib().cadre_gae(_lib.ptr(rewards), _lib.ptr(values), _lib.ptr(masks), _lib.ptr(next_value),
                                        _lib.ptr(returns), _lib.ptr(adv), E, T1 - 1, ctypes.c_float(gamma),
                                        ctypes.c_float(tau), int(bool(normalize)), _lib.stream_ptr(rewards.device)))
    return returns, adv
