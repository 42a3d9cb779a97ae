"""CadreAgent drop-in (reference: ppo_agent/agent.py:8-271).

Same constructor, attributes (`model_dict`, `vae_model`, `device`, `hidden_state`) and methods
(`act`, `get_value`, `update_policy`, `update_model`, `convert_action`, `avg_action`, `get_latent_feature`,
`save_snapshot`, `load_snapshot`). Every tensor computation runs in libcadre_sm100.so:
  get_latent_feature -> cadre_encoder_forward_u8 (uint8 ingest + DANet encoder + measurement concat)
  act / get_value    -> cadre_ppo_evaluate (routed LSTM + actor-critic forward)
  update_policy      -> cadre_ppo_update  (forward + hand-written backward; gradients land in the flat buffer)
Host code only marshals pointers and draws the action sample (torch.distributions.Categorical on the host,
distributions.py:96-99).
"""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import ppo as _ppo
from . import ppo_params
from ._lib import CadreError
from .models import create_model, get_vae_output, load_model_dict, save_model_dict
from .storage import MiniBatch


def _as_minibatch_views(samples, device):
    """Accept either a storage.MiniBatch or the reference's 9-tuple of tensors (storage.py:117-120)."""
    if isinstance(samples, MiniBatch):
        return samples.storage, samples.advantages, list(samples.indices)
    obs, action, value_preds, returns, masks, old_lp, adv, (hn, cn), command = samples
    mb = action.shape[0]
    seq = obs.shape[0] // mb
    st = SimpleNamespace(
        obs=obs.view(seq, mb, -1).permute(1, 0, 2).contiguous().to(device, torch.float32),
        action=action.contiguous().to(device, torch.int64), value_preds=value_preds.contiguous().to(device),
        returns=returns.contiguous().to(device), action_log_probs=old_lp.contiguous().to(device),
        hn=hn.contiguous().to(device), cn=cn.contiguous().to(device),
        command=command.contiguous().to(device, torch.int32))
    return st, adv.contiguous().to(device, torch.float32), list(range(mb))


class CadreAgent(object):
    def __init__(self, rank, model_cfg, frame, STEER_CONTROL, THROTTLE_CONTROL, ent_coeff, value_coeff, clip_coeff,
                 clip, danet_state=None, ppo_state=None, mini_batch=100, max_encoder_batch=64):
        self.rank = rank
        self.vae_model, self.model_dict = create_model(model_cfg, load_vae=True, danet_state=danet_state,
                                                       ppo_state=ppo_state, max_batch=max_encoder_batch)
        self.owner = self.model_dict.owner
        self.use_lstm = model_cfg["use_lstm"]
        if not self.use_lstm:
            raise CadreError("use_lstm=False is not on the hot path (agent_config.py:30 sets it True)")
        self.command_num = model_cfg["command_num"]
        self.device = torch.device("cuda:" + str(model_cfg["device_num"]))
        self.vae_device = torch.device("cuda:" + str(model_cfg["vae_device"]))
        self.STEER_CONTROL = STEER_CONTROL
        self.THROTTLE_CONTROL = THROTTLE_CONTROL
        self.ent_coeff, self.value_coeff, self.clip_coeff, self.clip = ent_coeff, value_coeff, clip_coeff, clip
        self.lstm_input, self.vae_params = get_vae_output(model_cfg)
        self.use_vae = True
        self.frame = frame
        self.pre_latent_feature = None
        # agent.py:38-40: zeros, and never updated by act()
        self.hidden_state = (torch.zeros(1, self.lstm_input, device=self.device),
                             torch.zeros(1, self.lstm_input, device=self.device))
        self._act_engine = _ppo.PpoEngine(1, 1, clip, value_coeff, clip_coeff, ent_coeff, self.device)
        self._upd_engine = None
        self._mini_batch = mini_batch
        self._zero_a = torch.zeros(1, 1, dtype=torch.int64, device=self.device)
        self._zero_f = torch.zeros(1, 1, device=self.device)

    # ------------------------------------------------------------------ perception
    def get_latent_feature(self, tick_data):
        """agent.py:97-112: [8,530] fp32 on self.device. rgb u8 [8,144,256,3], route_fig u8 [8,256,144],
        measurements f64 [8,3] (host numpy, as produced by the env wrapper, env_wrapper.py:900-914)."""
        rgb = torch.from_numpy(np.ascontiguousarray(tick_data["rgb"])).to(self.vae_device, non_blocking=True)
        route = torch.from_numpy(np.ascontiguousarray(tick_data["route_fig"])).to(self.vae_device, non_blocking=True)
        meas = torch.from_numpy(np.ascontiguousarray(tick_data["measurements"], dtype=np.float64)).to(
            self.vae_device, non_blocking=True)
        feat = self.vae_model.forward_u8(rgb, route, meas)
        return feat.to(self.device)

    # ------------------------------------------------------------------ acting
    def _evaluate_one(self, steer_obs, steer_cmd, throttle_obs, throttle_cmd, actions=None):
        sts = []
        for obs, cmd, h in ((steer_obs, steer_cmd, 0), (throttle_obs, throttle_cmd, 1)):
            a = self._zero_a if actions is None else actions[h].reshape(1, 1).to(self.device, torch.int64)
            sts.append(SimpleNamespace(
                obs=obs.reshape(1, 8, self.lstm_input).contiguous(), action=a, value_preds=self._zero_f,
                returns=self._zero_f, action_log_probs=self._zero_f, hn=self.hidden_state[0],
                cn=self.hidden_state[1],
                command=torch.full((1, 1), int(cmd), dtype=torch.int32, device=self.device)))
        out = self._act_engine.evaluate([tuple(sts)], [(self._zero_f, self._zero_f)], [[[0], [0]]],
                                        self.owner.params)
        return out  # [2, 1, 36]

    def act(self, tick_data):
        """agent.py:114-141."""
        command = tick_data["command"]
        ppo_feature = self.get_latent_feature(tick_data)
        out = self._evaluate_one(ppo_feature, command, ppo_feature, command).cpu()
        actions, log_probs, values = [], [], []
        for h, A in ((0, 33), (1, 3)):
            logits = out[h, 0, 3:3 + A]
            probs = F.softmax(logits, dim=-1).unsqueeze(0)                       # distributions.py:96-99
            action = torch.distributions.Categorical(probs=probs).sample()      # [1]
            actions.append(action[0])
            log_probs.append(logits[action[0]].reshape(1, 1).to(self.device))
            values.append(out[h, 0, 0].reshape(1, 1).to(self.device))
        return ppo_feature, actions, log_probs, values, self.hidden_state

    def get_value(self, done, steer_batch, throttle_batch):
        """agent.py:143-164."""
        if done:
            return torch.zeros(1), torch.zeros(1)
        (s_obs, s_cmd), (t_obs, t_cmd) = steer_batch, throttle_batch
        out = self._evaluate_one(s_obs.to(self.device), s_cmd, t_obs.to(self.device), t_cmd)
        return out[0, 0, 0].reshape(1, 1), out[1, 0, 0].reshape(1, 1)

    # ------------------------------------------------------------------ learning
    def update_policy(self, steer_samples, throttle_samples):
        """agent.py:166-237: returns (value_loss*coeff, action_loss*coeff, entropy*coeff) as Python floats;
        the gradient of the total loss is left in the flat gradient buffer (`model_dict[...]...grad`)."""
        s_st, s_adv, s_idx = _as_minibatch_views(steer_samples, self.device)
        t_st, t_adv, t_idx = _as_minibatch_views(throttle_samples, self.device)
        mb = len(s_idx)
        if self._upd_engine is None or self._upd_engine.mini_batch != mb:
            self._upd_engine = _ppo.PpoEngine(1, mb, self.clip, self.value_coeff, self.clip_coeff, self.ent_coeff,
                                              self.device)
        L = self._upd_engine.update([(s_st, t_st)], [(s_adv, t_adv)], [[s_idx, t_idx]], self.owner.params,
                                    self.owner.grads).sum(1)[0].cpu()
        return (L[0].item() * self.value_coeff, L[1].item() * self.clip_coeff, L[2].item() * self.ent_coeff)

    def update_model(self, shared_model_list):
        """agent.py:239-243: pull parameters from the shared model."""
        src = shared_model_list.owner if hasattr(shared_model_list, "owner") else None
        if src is not None:
            if src is not self.owner:
                self.owner.params.copy_(src.params)
            return
        for name in self.model_dict:
            self.model_dict[name].load_state_dict(shared_model_list[name].state_dict())

    # ------------------------------------------------------------------ misc (agent.py:77-95, 245-271)
    def convert_action(self, discrete_action):
        steer = self.STEER_CONTROL[discrete_action[0].item()]
        throttle, brake = self.THROTTLE_CONTROL[discrete_action[1].item()]
        return [steer, throttle, brake]

    def avg_action(self, discrete_action_list):
        n = len(discrete_action_list)
        control = np.array([self.convert_action(a) for a in discrete_action_list]).mean(0).tolist()
        if n > 1 and control[-1] < 0.5:
            control[-1] = 0.0
        return control

    def save_snapshot(self, model_path):
        """agent.py:245-260; all 16 modules (the reference forgets `throttle_lstm_*`), see models.save_model_dict."""
        save_model_dict(self.model_dict, model_path)

    def load_snapshot(self, model_path, device=None):
        """agent.py:262-271; reads this package's snapshots and the reference's pickled-module files."""
        load_model_dict(self.model_dict, model_path, device)

    def pre_process(self, tick_data):
        """agent.py:43-75 for callers that want the network input itself: fp32 [S,4,144,256] on `vae_device`.
        The hot path does NOT use this: `get_latent_feature` ingests the uint8 arrays directly (the same arithmetic is
        fused into the encoder's first kernel). Unlike the reference the caller's `route_fig` array is left
        unmodified (the in-place write-back of agent.py:51-54 is idempotent, so later ticks see the same values)."""
        return pre_process_tensors(torch.from_numpy(np.ascontiguousarray(tick_data["rgb"])).to(self.vae_device),
                                   torch.from_numpy(np.ascontiguousarray(tick_data["route_fig"])).to(self.vae_device))


def pre_process_tensors(rgb, route):
    """agent.py:43-75 on torch tensors of any device: rgb u8 [S,H,W,3] -> rgb/255 computed in float64 and cast to
    float32 (agent.py:46); route u8 [S,W,H] max-normalised per frame THROUGH uint8 (agent.py:51-54 writes the
    quotient back into the uint8 array, truncating it to {0,1}; frames whose max is 0 stay as they are), then
    transposed to [S,1,H,W] and concatenated as the fourth channel."""
    img = (rgb.double() / 255.0).float().permute(0, 3, 1, 2)
    mx = route.flatten(1).max(1).values.view(-1, 1, 1)
    norm = (route.double() / mx.double().clamp_min(1.0)).to(torch.uint8)     # float64 -> uint8 truncates
    route = torch.where(mx > 0, norm, route)
    return torch.cat([img, route.float().transpose(1, 2).unsqueeze(1)], 1).contiguous()


__all__ = ["CadreAgent", "ppo_params"]
