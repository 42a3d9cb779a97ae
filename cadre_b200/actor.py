"""Batched acting with a per-environment frame cache (reference: CadreAgent.act,
ppo_agent/agent.py:114-141, one environment per process).

The env wrapper hands every tick the last `seq_length` = 8 frames (env_wrapper.py:900-914) and the window slides by
ONE frame per tick, while the frozen perception encoder maps each frame independently (eval-mode BN, per-frame
route normalisation, agent.py:43-75). So, per environment, seven of the eight [530] features of a tick were
already computed for the previous tick. `BatchedActor` keeps them in a device ring, encodes only the newest frame of
each of E environments in one batch (E frames instead of 8E), and runs the per-command LSTM + actor-critic once for
all E environments (rows routed by command). Results equal E independent `CadreAgent.act` calls on the full windows;
only the order of the host-side Categorical draws differs (environment-major here)."""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import ppo as _ppo


class BatchedActor:
    def __init__(self, agent, num_envs, verify_window=True):
        """agent: a cadre_b200.agent.CadreAgent (owns the encoder and the flat PPO parameters)."""
        self.agent, self.E = agent, int(num_envs)
        self.device = agent.device
        self.F, self.S = agent.lstm_input, 8
        self.verify_window = verify_window
        self.feats = torch.zeros(self.E, self.S, self.F, device=self.device)
        # newest frame of the previous tick per env: the window slid by one iff it is now the second newest
        self._last = [None] * self.E
        self._engine = _ppo.PpoEngine(1, self.E, agent.clip, agent.value_coeff, agent.clip_coeff, agent.ent_coeff,
                                      self.device)
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=self.device)  # noqa: E731
        self._st = [SimpleNamespace(obs=self.feats, action=z(self.E, 1, dtype=torch.int64), value_preds=z(self.E, 1),
                                    returns=z(self.E, 1), action_log_probs=z(self.E, 1), hn=z(self.E, self.F),
                                    cn=z(self.E, self.F), command=z(self.E, 1, dtype=torch.int32)) for _ in range(2)]
        self._adv = (z(self.E, 1), z(self.E, 1))
        self._idx = [[list(range(self.E)), list(range(self.E))]]
        # pinned staging for the newest frames (one H2D per tensor and tick)
        self._rgb_h = torch.empty(self.E, 144, 256, 3, dtype=torch.uint8, pin_memory=True)
        self._route_h = torch.empty(self.E, 256, 144, dtype=torch.uint8, pin_memory=True)
        self._meas_h = torch.empty(self.E, 3, dtype=torch.float64, pin_memory=True)
        self._cmd_h = torch.empty(self.E, 1, dtype=torch.int32, pin_memory=True)
        self.frames_encoded = 0

    # ------------------------------------------------------------------ perception with the frame cache
    def _slid_by_one(self, e, tick):
        last = self._last[e]
        if last is None:
            return False
        if not self.verify_window:
            return True
        return (np.array_equal(tick["rgb"][-2], last[0]) and np.array_equal(tick["route_fig"][-2], last[1])
                and np.array_equal(tick["measurements"][-2], last[2]))

    def encode(self, ticks):
        """ticks: list of E tick_data dicts. Updates and returns the [E, 8, 530] feature windows."""
        assert len(ticks) == self.E
        full = [e for e in range(self.E) if not self._slid_by_one(e, ticks[e])]
        enc = self.agent.vae_model
        dev = self.agent.vae_device
        if len(full) < self.E:   # newest frame of every environment whose window slid by one
            inc = [e for e in range(self.E) if e not in set(full)] if full else list(range(self.E))
            n = len(inc)
            rgb_np, route_np, meas_np = self._rgb_h.numpy(), self._route_h.numpy(), self._meas_h.numpy()
            for i, e in enumerate(inc):
                np.copyto(rgb_np[i], ticks[e]["rgb"][-1])
                np.copyto(route_np[i], ticks[e]["route_fig"][-1])
                meas_np[i] = ticks[e]["measurements"][-1]
            new = enc.forward_u8(self._rgb_h[:n].to(dev, non_blocking=True), self._route_h[:n].to(dev, non_blocking=True),
                                 self._meas_h[:n].to(dev, non_blocking=True)).to(self.device)
            if n == self.E:
                self.feats[:, :-1] = self.feats[:, 1:].clone()
                self.feats[:, -1] = new
            else:
                ids = torch.as_tensor(inc, device=self.device)
                self.feats[ids, :-1] = self.feats[ids, 1:].clone()
                self.feats[ids, -1] = new
            self.frames_encoded += n
        for e in full:   # first tick after a reset (or a window that did not slide): all eight frames
            t = ticks[e]
            self.feats[e] = enc.forward_u8(
                torch.from_numpy(np.ascontiguousarray(t["rgb"])).to(dev),
                torch.from_numpy(np.ascontiguousarray(t["route_fig"])).to(dev),
                torch.from_numpy(np.ascontiguousarray(t["measurements"], dtype=np.float64)).to(dev)).to(self.device)
            self.frames_encoded += self.S
        for e in range(self.E):
            t = ticks[e]
            self._last[e] = (np.array(t["rgb"][-1], copy=True), np.array(t["route_fig"][-1], copy=True),
                             np.array(t["measurements"][-1], copy=True))
        return self.feats

    def reset(self, env_id):
        """Forget environment `env_id`'s window (call after env.reset())."""
        self._last[env_id] = None

    # ------------------------------------------------------------------ acting
    def act_batch(self, ticks):
        """One tick of all E environments. Returns (features [E,8,530], actions int64 [E,2], log_probs [E,2],
        values [E,2]) on the device; column 0 = steer, 1 = throttle. One Categorical draw per head for the whole
        batch (distributions.py:96-99 semantics, environment-major draw order)."""
        feats = self.encode(ticks)
        for e, t in enumerate(ticks):
            self._cmd_h[e, 0] = int(t["command"])
        cmd = self._cmd_h.to(self.device, non_blocking=True)
        for st in self._st:
            st.command.copy_(cmd)
        out = self._engine.evaluate([tuple(self._st)], [self._adv], self._idx, self.agent.owner.params)  # [2, E, 36]
        actions, log_probs = [], []
        for h, A in ((0, 33), (1, 3)):
            logits = out[h, :, 3:3 + A]                                   # normalised logits = log-probabilities
            a = torch.multinomial(F.softmax(logits, dim=-1), 1)           # [E,1]
            actions.append(a)
            log_probs.append(logits.gather(1, a))
        return feats, torch.cat(actions, 1), torch.cat(log_probs, 1), out[:, :, 0].t().contiguous()

    def act(self, ticks):
        """Reference-shaped results: per environment the 5-tuple of CadreAgent.act: (feature [8,530],
        [a_steer, a_throttle] 0-dim int64, [lp_s, lp_t] [1,1], [v_s, v_t] [1,1], hidden_state)."""
        feats, actions, log_probs, values = self.act_batch(ticks)
        a_cpu = actions.cpu()
        return [(feats[e].clone(), [a_cpu[e, 0], a_cpu[e, 1]],
                 [log_probs[e, 0].reshape(1, 1), log_probs[e, 1].reshape(1, 1)],
                 [values[e, 0].reshape(1, 1), values[e, 1].reshape(1, 1)], self.agent.hidden_state)
                for e in range(self.E)]
