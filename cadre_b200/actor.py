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
        self._last = [None] * self.E          # previous tick's host arrays per env (for the sliding-window check)
        self._engine = _ppo.PpoEngine(1, self.E, agent.clip, agent.value_coeff, agent.clip_coeff, agent.ent_coeff,
                                      self.device)
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=self.device)  # noqa: E731
        self._st = [SimpleNamespace(obs=self.feats, action=z(self.E, 1, dtype=torch.int64), value_preds=z(self.E, 1),
                                    returns=z(self.E, 1), action_log_probs=z(self.E, 1), hn=z(self.E, self.F),
                                    cn=z(self.E, self.F), command=z(self.E, 1, dtype=torch.int32)) for _ in range(2)]
        self._adv = (z(self.E, 1), z(self.E, 1))
        self._idx = [[list(range(self.E)), list(range(self.E))]]
        self.frames_encoded = 0

    # ------------------------------------------------------------------ perception with the frame cache
    def _slid_by_one(self, e, tick):
        last = self._last[e]
        if last is None:
            return False
        if not self.verify_window:
            return True
        return all(np.array_equal(tick[k][:-1], last[k][1:]) for k in ("rgb", "route_fig", "measurements"))

    def encode(self, ticks):
        """ticks: list of E tick_data dicts. Updates and returns the [E, 8, 530] feature windows."""
        assert len(ticks) == self.E
        full = [e for e in range(self.E) if not self._slid_by_one(e, ticks[e])]
        inc = [e for e in range(self.E) if e not in full]
        enc = self.agent.vae_model
        if inc:   # newest frame of every environment whose window slid by one
            rgb = np.stack([ticks[e]["rgb"][-1] for e in inc])
            route = np.stack([ticks[e]["route_fig"][-1] for e in inc])
            meas = np.stack([np.asarray(ticks[e]["measurements"][-1], dtype=np.float64) for e in inc])
            new = self._encode_host(enc, rgb, route, meas)
            ids = torch.as_tensor(inc, device=self.device)
            self.feats[ids, :-1] = self.feats[ids, 1:].clone()
            self.feats[ids, -1] = new
            self.frames_encoded += len(inc)
        for e in full:   # first tick after a reset (or a window that did not slide): all eight frames
            t = ticks[e]
            self.feats[e] = self._encode_host(enc, np.asarray(t["rgb"]), np.asarray(t["route_fig"]),
                                              np.asarray(t["measurements"], dtype=np.float64))
            self.frames_encoded += self.S
        for e in range(self.E):
            t = ticks[e]
            self._last[e] = {k: np.array(t[k], copy=True) for k in ("rgb", "route_fig", "measurements")}
        return self.feats

    def _encode_host(self, enc, rgb, route, meas):
        dev = self.agent.vae_device
        return enc.forward_u8(torch.from_numpy(np.ascontiguousarray(rgb)).to(dev),
                              torch.from_numpy(np.ascontiguousarray(route)).to(dev),
                              torch.from_numpy(np.ascontiguousarray(meas)).to(dev)).to(self.device)

    def reset(self, env_id):
        """Forget environment `env_id`'s window (call after env.reset())."""
        self._last[env_id] = None

    # ------------------------------------------------------------------ acting
    def act(self, ticks):
        """Returns, per environment, the 5-tuple of CadreAgent.act: (feature [8,530], [a_steer, a_throttle] 0-dim
        int64, [lp_s, lp_t] [1,1], [v_s, v_t] [1,1], hidden_state)."""
        feats = self.encode(ticks)
        cmd = torch.tensor([[int(t["command"])] for t in ticks], dtype=torch.int32, device=self.device)
        for st in self._st:
            st.command.copy_(cmd)
        out = self._engine.evaluate([tuple(self._st)], [self._adv], self._idx, self.agent.owner.params).cpu()
        results = []
        for e in range(self.E):
            actions, log_probs, values = [], [], []
            for h, A in ((0, 33), (1, 3)):
                logits = out[h, e, 3:3 + A]
                probs = F.softmax(logits, dim=-1).unsqueeze(0)                     # distributions.py:96-99
                action = torch.distributions.Categorical(probs=probs).sample()
                actions.append(action[0])
                log_probs.append(logits[action[0]].reshape(1, 1).to(self.device))
                values.append(out[h, e, 0].reshape(1, 1).to(self.device))
            results.append((feats[e].clone(), actions, log_probs, values, self.agent.hidden_state))
        return results
