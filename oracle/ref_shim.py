"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference modules from /root/reference so that
`oracle/restate.py` can be validated against them and `oracle/make_golden.py` can write golden vectors.
/root/reference exists only in the build container; callers must check `available()` first. Nothing in
`-m gpu` tests, `smoke()` or `bench.py` may depend on this module at run time.
"""
import importlib.util
import os
import sys
import tempfile
import types

import torch

REF = os.environ.get("CADRE_REF", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "ppo_agent"))


class AttrDict(dict):
    """Minimal stand-in for the mmcv-style Config (ppo_agent/meta/config.py needs addict + yapf, absent here)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _to_attr(d):
    if isinstance(d, dict) and all(isinstance(k, str) for k in d):
        return AttrDict({k: _to_attr(v) for k, v in d.items()})
    return d


_installed = False


def install(scratch=None):
    """Put the reference on sys.path behind the two import shims; returns the CHALLENGE_DIR scratch path."""
    global _installed
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    if scratch is None:
        scratch = os.environ.get("CHALLENGE_DIR") or tempfile.mkdtemp(prefix="cadre_ref_")
    os.environ["CHALLENGE_DIR"] = scratch
    if not _installed:
        sys.path[:0] = [REF, REF + "/carla_perception"]
        sys.modules.setdefault("torchsnooper", types.ModuleType("torchsnooper"))
        pkg = types.ModuleType("carla_perception")
        pkg.__path__ = [REF + "/carla_perception"]
        sys.modules["carla_perception"] = pkg  # skip carla_perception/__init__.py (tensorboardX, skimage)
        _installed = True
    return scratch


def load_agent_config():
    """config_files/agent_config.py as attribute dicts, with device_num = vae_device = -1 (CPU)."""
    spec = importlib.util.spec_from_file_location("cadre_ref_agent_config", REF + "/config_files/agent_config.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = AttrDict(rollout_cfg=_to_attr(mod.rollout_cfg), agent_cfg=_to_attr(mod.agent_cfg),
                   train_cfg=_to_attr(mod.train_cfg), env_cfg=_to_attr(mod.env_cfg))
    cfg.agent_cfg.model_cfg.device_num = -1
    cfg.agent_cfg.model_cfg.vae_device = -1
    return cfg


def build_reference_danet(state):
    """DANet (carla_perception/Networks/danet.py:73) in eval mode with the on-path fixture tensors loaded."""
    install()
    from carla_perception.Config.auto_danet import danet_config
    from carla_perception.Networks.danet import DANet
    cfg = danet_config()
    net = DANet(cfg.networks["autoencoder"])
    missing, unexpected = net.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    on_path = ("backbone.", "da_head.", "visual_conv.", "bc_conv.", "inter_task_att.")
    assert not [k for k in missing if k.startswith(on_path)], "fixture misses on-path keys"
    net.eval()
    return net, cfg


def build_reference_agent(danet_state, ppo_state):
    """CadreAgent (ppo_agent/agent.py:8) on CPU with fixture weights: writes the fake pretrained checkpoint
    that create_model(load_vae=True) loads (models.py:54-63), then overwrites the 16 PPO modules."""
    scratch = install()
    net, cfg = build_reference_danet(danet_state)
    path = cfg.networks["autoencoder"]["pretrained_path"]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({"autoencoder": net.state_dict()}, path)
    from ppo_agent.agent import CadreAgent
    acfg = load_agent_config().agent_cfg
    agent = CadreAgent(**acfg)
    for name, sd in ppo_state.items():
        agent.model_dict[name].load_state_dict(sd, strict=True)
    return agent, scratch


def reference_storage(st, rollout_cfg=None):
    """RolloutStorage (ppo_agent/storage.py:5) filled from a dict of tensors."""
    install()
    from ppo_agent.storage import RolloutStorage
    cfg = dict(load_agent_config().rollout_cfg) if rollout_cfg is None else dict(rollout_cfg)
    cfg["hidden_size"] = cfg["feature_dims"]
    cfg["num_steps"] = st["rewards"].shape[0] - 1
    rs = RolloutStorage(**cfg)
    for k, v in st.items():
        getattr(rs, k).copy_(v)
    return rs
