"""Host-side mirror of ppo_agent/models.py for the hot path: `create_model`, `get_vae_output`,
`Shared_grad_buffers`, and `FlatModule`, a module-like window onto the flat parameter buffer.

The 16 reference modules (8 x Model, 8 x LSTM; models.py:44-126) do not exist as nn.Modules here: all 128
tensors live in ONE flat fp32 buffer (cadre_b200/ppo_params.py) that the CUDA kernels consume directly.
`model_dict[name]` keeps the reference's naming and the `named_parameters / state_dict / load_state_dict /
zero_grad` surface that `train.py`, `chief.py` and snapshot code touch, implemented as copies in / out of
the flat buffer.
"""
import math
from collections import OrderedDict

import torch

from . import ppo_params

Z_DIMS = 256            # carla_perception/Config/auto_danet.py:145
LATENT_CHANNELS = 4     # RGB + route (auto_danet.py:111-119)


def get_vae_output(model_cfg):
    """models.py:33-42: observation width = 2*z_dims (+ measurement_dim) for the CoPM encoders."""
    vae_params_cfg = model_cfg["vae_params"]
    measurement_dim = model_cfg["measurement_dim"]
    if vae_params_cfg in ("CoPM", "CoPM w/o att"):
        obs_dim = 2 * Z_DIMS + measurement_dim
    else:
        obs_dim = Z_DIMS + measurement_dim
    return obs_dim, None


class FlatParameter:
    """One reference tensor inside the flat buffers: `.data` / `.grad` are reference-shaped COPIES."""

    def __init__(self, owner, module, name):
        self._o, self._m, self._n = owner, module, name
        self.requires_grad = True

    def _get(self, flat):
        view, perm = ppo_params.tensor_view(flat, self._m, self._n)
        return (ppo_params._gate_deinterleave(view) if perm else view).clone()

    @property
    def data(self):
        return self._get(self._o.params)

    @property
    def grad(self):
        return self._get(self._o.grads)

    def size(self):
        return self.data.size()


class FlatModule:
    """Stand-in for one `Model` / `LSTM` nn.Module (models.py:130-212) backed by a FlatParams owner."""

    def __init__(self, owner, name):
        self._o, self.name = owner, name

    def named_parameters(self):
        for pn in ppo_params.module_param_names(self.name):
            yield pn, FlatParameter(self._o, self.name, pn)

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def state_dict(self):
        return OrderedDict((pn, p.data) for pn, p in self.named_parameters())

    def load_state_dict(self, sd, strict=True):
        for pn in ppo_params.module_param_names(self.name):
            view, perm = ppo_params.tensor_view(self._o.params, self.name, pn)
            t = sd[pn].detach().to(device=view.device, dtype=torch.float32)
            view.copy_(ppo_params._gate_interleave(t) if perm else t)

    def zero_grad(self):
        for pn in ppo_params.module_param_names(self.name):
            view, _ = ppo_params.tensor_view(self._o.grads, self.name, pn)
            view.zero_()

    def train(self):
        return self

    def eval(self):
        return self

    def to(self, device):
        return self

    def share_memory(self):
        return self


class FlatParams:
    """Owner of the flat parameter / gradient buffers (+ Adam moments once an optimizer attaches)."""

    def __init__(self, device, state=None):
        self.device = torch.device(device)
        self.params = torch.zeros(ppo_params.TOTAL, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.model_dict = OrderedDict((m, FlatModule(self, m)) for m in ppo_params.MODULE_ORDER)
        if state is not None:
            self.load_state(state)

    def load_state(self, state):
        self.params.copy_(ppo_params.pack_state(state, self.device))

    def state(self):
        return ppo_params.unpack_state(self.params)


def _orthogonal(shape, gain, gen):
    t = torch.empty(shape)
    torch.nn.init.orthogonal_(t, gain=gain, generator=gen)
    return t


def init_state(seed=None):
    """Fresh parameters with the reference's initialisers: orthogonal LSTM weights, zero biases
    (models.py:133-137); orthogonal actor (gain 0.01) / critic (gain 1) linears, zero biases
    (distributions.py:29-40, models.py:165-177). Draws from its own generator (the reference uses the global
    RNG in module-construction order, which is not reproduced here)."""
    gen = torch.Generator()
    gen.manual_seed(torch.seed() if seed is None else seed)
    F, H = ppo_params.F, ppo_params.HID
    state = {}
    for m in ppo_params.MODULE_ORDER:
        d = {}
        if "_lstm_" in m:
            d["rnn.weight_ih"] = _orthogonal((4 * F, F), 1.0, gen)
            d["rnn.weight_hh"] = _orthogonal((4 * F, F), 1.0, gen)
            d["rnn.bias_ih"] = torch.zeros(4 * F)
            d["rnn.bias_hh"] = torch.zeros(4 * F)
        else:
            A = ppo_params.ACTIONS[m.split("_")[0]]
            for prefix, gain, out in (("control.linear.", 0.01, A), ("critic.", 1.0, 1)):
                d[prefix + "0.weight"] = _orthogonal((H, F), gain, gen)
                d[prefix + "0.bias"] = torch.zeros(H)
                d[prefix + "2.weight"] = _orthogonal((H, H), gain, gen)
                d[prefix + "2.bias"] = torch.zeros(H)
                d[prefix + "4.weight"] = _orthogonal((out, H), gain, gen)
                d[prefix + "4.bias"] = torch.zeros(out)
            d = {k: d[k] for k in ppo_params.module_param_names(m)}
        state[m] = d
    return state


def create_model(model_cfg, load_vae=False, danet_state=None, ppo_state=None, max_batch=64):
    """models.py:44-126. Returns (vae_model, model_dict): `vae_model` is a `cadre_b200.encoder.Encoder` (or None
    when load_vae is False, as in main.py:38); `model_dict` maps the 16 reference module names to FlatModule
    windows of one shared `FlatParams` (reachable as `model_dict.owner`).

    `danet_state`: the `'autoencoder'` state dict of the reference checkpoint (models.py:55-63). When omitted
    the checkpoint at `model_cfg['pretrained_path']` is read with torch.load (same file format)."""
    dev = model_cfg["device_num"]
    if dev == -1:
        from ._lib import CadreError
        raise CadreError("device_num = -1 (CPU) is not supported: cadre_b200 has no CPU path")
    device = torch.device("cuda:" + str(dev))
    vae_model = None
    if load_vae:
        from .encoder import Encoder
        if danet_state is None:
            ckpt = torch.load(model_cfg["pretrained_path"], map_location="cpu")
            danet_state = ckpt["autoencoder"]
        vae_dev = torch.device("cuda:" + str(model_cfg["vae_device"]))
        vae_model = Encoder(danet_state, vae_dev, max_batch=max_batch)
    owner = FlatParams(device, ppo_state if ppo_state is not None else init_state())
    model_dict = ModelDict(owner)
    return vae_model, model_dict


class ModelDict(OrderedDict):
    """model_dict of the reference plus a handle on the shared flat buffers."""

    def __init__(self, owner):
        super().__init__(owner.model_dict)
        self.owner = owner


class Shared_grad_buffers:
    """models.py:219-258: running SUM of worker gradients (never averaged — `average_gradient` is unused in the
    reference). Here one flat buffer; across GPUs the sum is the NCCL all-reduce in cadre_b200.learner."""

    def __init__(self, model_list, device=None):
        self.owner = model_list.owner
        self.grads = torch.zeros_like(self.owner.params)
        self.count = 0

    def add_gradient(self, model_list):
        self.grads += model_list.owner.grads
        self.count += 1

    def reset(self):
        self.count = 0
        self.grads.zero_()
