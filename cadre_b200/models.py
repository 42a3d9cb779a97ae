"""Host-side mirror of ppo_agent/models.py for the hot path: `create_model`, `get_vae_output`,
`Shared_grad_buffers`, and `FlatModule`, a module-like window onto the flat parameter buffer.

The 16 reference modules (8 x Model, 8 x LSTM; models.py:44-126) do not exist as nn.Modules here: all 128
tensors live in ONE flat fp32 buffer (cadre_b200/ppo_params.py) that the CUDA kernels consume directly.
`model_dict[name]` keeps the reference's naming and the `named_parameters / state_dict / load_state_dict /
zero_grad` surface that `train.py`, `chief.py` and snapshot code touch, implemented as copies in / out of
the flat buffer.
"""
import math
import os
from collections import OrderedDict

import torch

from . import ppo_params
from ._lib import CadreError

Z_DIMS = 256            # carla_perception/Config/auto_danet.py:145
LATENT_CHANNELS = 4     # RGB + route (auto_danet.py:111-119)

# carla_perception/Config/auto_danet.py:14,24,41,46-56: load_epoch 90, input_mode 9, output_mode 12, version "danet",
# train_town "nocrash_IL_n10_k1234_r40", exp_suffix "34"
_LOAD_EPOCH = 90
_EXP_SUFFIX = "34"
_EXP_DIR = "danet912_nocrash_IL_n10_k1234_r40"
# tensors of the checkpoint that `get_latent_feature(x, "concate")` reads (danet.py:216-238); everything else in the
# file (decoder heads, speed fc, ...) is off the path and ignored
ON_PATH_PREFIXES = ("backbone.", "da_head.", "visual_conv.", "bc_conv.", "inter_task_att.")


class DanetParams:
    """The two fields of `danet_config()` (auto_danet.py:7-171) that the PPO side reads through
    `get_vae_output(...)[1]`: `.networks['autoencoder']['z_dims']` and `['pretrained_path']`."""

    def __init__(self, pretrained_path=None):
        self.load_epoch = _LOAD_EPOCH
        self.networks = {"autoencoder": {"z_dims": Z_DIMS, "pretrained": True, "pretrained_path": pretrained_path}}


def default_pretrained_path():
    """auto_danet.py:161-171: $CHALLENGE_DIR/carla_perception/Experiments34/danet912_nocrash_IL_n10_k1234_r40/
    net_epoch90. Like the reference, a missing CHALLENGE_DIR is a KeyError."""
    return os.path.join(os.environ["CHALLENGE_DIR"], "carla_perception/Experiments" + _EXP_SUFFIX, _EXP_DIR,
                        "net_epoch" + str(_LOAD_EPOCH))


def get_vae_output(model_cfg):
    """models.py:33-42: observation width = 2*z_dims (+ measurement_dim) for the CoPM encoders, and the
    perception config. `model_cfg['pretrained_path']` (an extension: the reference config has no such key)
    overrides the $CHALLENGE_DIR-derived checkpoint path."""
    vae_params_cfg = model_cfg["vae_params"]
    measurement_dim = model_cfg["measurement_dim"]
    path = model_cfg.get("pretrained_path") if hasattr(model_cfg, "get") else None
    if path is None and "CHALLENGE_DIR" in os.environ:
        path = default_pretrained_path()
    vae_params = DanetParams(path)
    if vae_params_cfg in ("CoPM", "CoPM w/o att"):
        obs_dim = 2 * Z_DIMS + measurement_dim
    else:
        obs_dim = Z_DIMS + measurement_dim
    return obs_dim, vae_params


def load_danet_checkpoint(path):
    """models.py:55-63: the perception checkpoint is `{'epoch', 'metric', 'autoencoder': state_dict}` written by
    experiments_builder.py:442-...; returns the `'autoencoder'` state dict restricted to the on-path tensors.
    Unlike the reference (which prints 'VAE model load fail' and keeps RANDOM weights when keys do not match,
    models.py:68-74) a checkpoint that lacks an on-path tensor is an error."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if "autoencoder" not in ckpt:
        raise CadreError(f"{path}: no 'autoencoder' entry (keys: {list(ckpt.keys())[:8]})")
    sd = ckpt["autoencoder"]
    return {k: v for k, v in sd.items() if k.startswith(ON_PATH_PREFIXES)}


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

    `danet_state`: the `'autoencoder'` state dict of the reference checkpoint (models.py:55-63), for callers that
    hold the weights in memory. When omitted the checkpoint is read from the path the reference uses:
    `danet_config().networks['autoencoder']['pretrained_path']` = $CHALLENGE_DIR/carla_perception/Experiments34/
    danet912_nocrash_IL_n10_k1234_r40/net_epoch90 (auto_danet.py:161-171), so `CadreAgent(**agent_cfg)` works
    with the reference's unmodified config."""
    dev = model_cfg["device_num"]
    if dev == -1:
        raise CadreError("device_num = -1 (CPU) is not supported: cadre_b200 has no CPU path")
    device = torch.device("cuda:" + str(dev))
    vae_model = None
    if load_vae:
        from .encoder import Encoder
        if model_cfg["vae_device"] == -1:
            raise CadreError("vae_device = -1 (CPU) is not supported: cadre_b200 has no CPU path")
        if danet_state is None:
            _, vae_params = get_vae_output(model_cfg)
            path = vae_params.networks["autoencoder"]["pretrained_path"]
            if path is None:
                path = default_pretrained_path()      # KeyError('CHALLENGE_DIR') like auto_danet.py:168
            danet_state = load_danet_checkpoint(path)
        vae_dev = torch.device("cuda:" + str(model_cfg["vae_device"]))
        vae_model = Encoder(danet_state, vae_dev, max_batch=max_batch)
    owner = FlatParams(device, ppo_state if ppo_state is not None else init_state())
    model_dict = ModelDict(owner)
    return vae_model, model_dict


class ModuleState:
    """Picklable snapshot entry: holds one module's state dict and answers `.state_dict()`, which is all the
    reference's `load_snapshot` asks of an entry (agent.py:266-267: `model_dict[name].state_dict()`)."""

    def __init__(self, sd=()):
        self.sd = OrderedDict(sd)

    def state_dict(self):
        return self.sd


# torch >= 2.6 loads with weights_only=True by default (the reference's torch.load call, agent.py:266, has no such
# argument): allow-list the class so that those loaders accept our snapshots as well
if hasattr(torch.serialization, "add_safe_globals"):
    torch.serialization.add_safe_globals([ModuleState])


def save_model_dict(model_dict, model_path):
    """agent.py:245-260 `save_snapshot`. The reference pickles the nn.Modules themselves and forgets the four
    `throttle_lstm_*` modules (it stores `steer_ppo_*` twice); here all 16 entries are written as `ModuleState`
    objects: readable by this package, and by the reference's `load_snapshot` whenever `cadre_b200` is importable
    in that process (it only calls `.state_dict()` on each entry)."""
    torch.save({name: ModuleState((k, v.cpu()) for k, v in m.state_dict().items()) for name, m in model_dict.items()},
               model_path)


def load_model_dict(model_dict, model_path, device=None):
    """agent.py:262-271 `load_snapshot`: accepts snapshots written by `save_model_dict`, plain dicts of state
    dicts, and the reference's own files (pickled `Model` / `LSTM` nn.Modules: unpickling those needs the
    reference's `ppo_agent.models` importable, as in any process that ran the reference). `device` is accepted for
    signature compatibility; tensors are copied into the flat parameter buffer wherever that lives.
    Errors are re-raised as ImportError like the reference."""
    try:
        snap = torch.load(model_path, map_location="cpu", weights_only=False)
        for name in snap:
            entry = snap[name]
            sd = entry.state_dict() if hasattr(entry, "state_dict") else entry
            model_dict[name].load_state_dict(sd)
    except Exception as e:  # agent.py:270-271
        raise ImportError("load snapshot error due to {}".format(e))


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
