"""Attribute-dict loader for agent_config.py-style Python config files (the reference uses an mmcv-style
`Config.fromfile`, ppo_agent/meta/config.py:239-247, which needs addict + yapf; the files themselves are plain
Python dict literals, so a 20-line loader is enough)."""
import importlib.util
import os


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _wrap(d):
    if isinstance(d, dict) and all(isinstance(k, str) for k in d):
        return AttrDict({k: _wrap(v) for k, v in d.items()})
    return d  # STEER_CONTROL / THROTTLE_CONTROL keep their int keys


DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_files", "agent_config.py")


def load_config(path=DEFAULT_CONFIG):
    spec = importlib.util.spec_from_file_location("cadre_agent_config", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return AttrDict({k: _wrap(getattr(mod, k)) for k in ("rollout_cfg", "agent_cfg", "train_cfg", "env_cfg")
                     if hasattr(mod, k)})
