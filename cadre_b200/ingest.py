"""Rollout ingest: uint8 camera frames in pinned HOST memory -> encoder features in the rollout storages' `obs`
(the step immediately upstream of the learner path: ppo_agent/agent.py:97-112 +
train.py:66-72).

What the reference does per environment tick: `act()` copies the 8-frame window to the GPU, pre-processes and encodes
all 8 frames, and `insert()` stores the [8,530] feature window in both heads' storages. A rollout of T ticks therefore
moves and encodes 8T frames per worker, although the window slides by ONE frame per tick (env_wrapper.py:900-914):
only T + 7 frames are distinct, and the frozen encoder maps frames independently (eval-mode BatchNorm, per-frame route
normalisation), so feature(window frame) is a function of the frame alone.

`RolloutIngest.encode(..., unique=True)` (the product path) ships each DISTINCT frame once over PCIe, encodes it
once and lets `cadre_window_scatter` assemble the sliding windows in `obs` (bit-identical features: same kernels, row
independent). `unique=False` ships and encodes every window frame like the reference's act() loop does (kept for
like-for-like measurements against round 1). In both modes the host-to-device copies run on a copy stream into
double-buffered staging tensors and overlap the encoder, whose chunks alternate between `streams` encoder instances on
their own CUDA streams (the tail of one chunk's persistent kernels is filled by the other chunk's CTAs); the first
chunks are small because the encoder cannot start before its first chunk has landed. `prefetch()` software-pipelines
across rollouts: the next rollout's distinct frames cross PCIe while the current rollout's update phase runs.
"""
import ctypes
import os

import torch

from . import _lib
from .encoder import Encoder


def _default_ramp():
    v = os.environ.get("CADRE_INGEST_RAMP")
    # geometric: each chunk's copy lands about when the previous chunk's encode ends (measured on B200, cfg 3's 828
    # distinct frames, tools/gpu_probe_ingest_ramp.py: 3.74 ms with (32, 96, 288) against 4.09 ms with round 1's
    # (128, 512); 2.63 ms with the frames already on the device)
    return tuple(int(x) for x in v.split(",") if x) if v is not None else (32, 96, 288)


def chunk_schedule(n, max_chunk, ramp=None):
    """Chunk sizes covering n frames: a short ramp (PCIe latency hiding), then `max_chunk`, then the remainder."""
    sizes, left = [], n
    for r in (_default_ramp() if ramp is None else ramp):
        if left > max_chunk and r < max_chunk:
            sizes.append(r)
            left -= r
    while left > 0:
        sizes.append(min(max_chunk, left))
        left -= sizes[-1]
    return sizes


class RolloutIngest:
    def __init__(self, danet_state, device, workers, num_steps, seq_length=8, feature_dims=530, max_chunk=640,
                 streams=2, encoders=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CadreError("RolloutIngest needs a CUDA device: there is no CPU fallback")
        self.W, self.T, self.S, self.F = int(workers), int(num_steps), int(seq_length), int(feature_dims)
        self.K = self.T + self.S - 1                      # distinct frames per worker and rollout
        self.max_chunk = int(max_chunk)
        self.NS = max(1, int(streams))
        self.encoders = list(encoders) if encoders is not None else [
            Encoder(danet_state, self.device, max_batch=self.max_chunk) for _ in range(self.NS)]
        assert len(self.encoders) == self.NS and all(e.max_batch >= self.max_chunk for e in self.encoders)
        self._lib = _lib.lib()
        with torch.cuda.device(self.device):
            self.enc_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.NS)]
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.NB = 2 * self.NS                         # staging buffers: two per encoder stream
            u8 = dict(dtype=torch.uint8, device=self.device)
            self.stage = [(torch.empty(self.max_chunk, 144, 256, 3, **u8), torch.empty(self.max_chunk, 256, 144, **u8),
                           torch.empty(self.max_chunk, 3, dtype=torch.float64, device=self.device))
                          for _ in range(self.NB)]
            self.staged_ev = [torch.cuda.Event() for _ in range(self.NB)]
            self.free_ev = [torch.cuda.Event() for _ in range(self.NB)]
            self.fork_ev = torch.cuda.Event()
            self.join_ev = [torch.cuda.Event() for _ in range(self.NS)]
            self._feats = {}
            self._landing = None                          # device copies of a prefetched rollout's distinct frames
            self._landing_key = None
            self._landed_ev = torch.cuda.Event()
            self._landing_free_ev = torch.cuda.Event()
            self._landing_free_ev.record(torch.cuda.current_stream(self.device))
        self.h2d_bytes_last = 0
        self.frames_encoded_last = 0

    def _feat_buffer(self, n):
        if n not in self._feats:
            self._feats[n] = torch.empty(n, self.F, device=self.device, dtype=torch.float32)
        return self._feats[n]

    # ------------------------------------------------------------------ host -> staging -> encoder
    def _encode_stream(self, rgb, route, meas, feats, host):
        """Encode n frames (flat leading dimension) into feats [n, F]. host=True: inputs are pinned host tensors,
        copied chunk by chunk on the copy stream; host=False: inputs are already on the device."""
        n = rgb.shape[0]
        sizes = chunk_schedule(n, self.max_chunk) if host else chunk_schedule(n, self.max_chunk, ramp=())
        main = torch.cuda.current_stream(self.device)
        self.fork_ev.record(main)
        for st in self.enc_streams:
            st.wait_event(self.fork_ev)
        if host:
            self.copy_stream.wait_event(self.fork_ev)   # previous consumers of the staging buffers are done
        s = 0
        for i, m in enumerate(sizes):
            st = self.enc_streams[i % self.NS]
            enc = self.encoders[i % self.NS]
            if host:
                k = i % self.NB
                with torch.cuda.stream(self.copy_stream):
                    if i >= self.NB:
                        self.copy_stream.wait_event(self.free_ev[k])
                    self.stage[k][0][:m].copy_(rgb[s:s + m], non_blocking=True)
                    self.stage[k][1][:m].copy_(route[s:s + m], non_blocking=True)
                    self.stage[k][2][:m].copy_(meas[s:s + m], non_blocking=True)
                    self.staged_ev[k].record(self.copy_stream)
                st.wait_event(self.staged_ev[k])
                a = (self.stage[k][0][:m], self.stage[k][1][:m], self.stage[k][2][:m])
            else:
                a = (rgb[s:s + m], route[s:s + m], meas[s:s + m])
            with torch.cuda.stream(st):
                enc.forward_u8(a[0], a[1], a[2], feats[s:s + m])
                if host:
                    self.free_ev[i % self.NB].record(st)
            s += m
        for k, st in enumerate(self.enc_streams):
            self.join_ev[k].record(st)
            main.wait_event(self.join_ev[k])
        if host:
            self.h2d_bytes_last = int(rgb[0].numel() + route[0].numel() + meas[0].numel() * 8) * n
        self.frames_encoded_last = n

    def _scatter(self, feats, obs, seq, steps):
        """cadre_window_scatter: feats [W][steps + seq - 1][F] -> obs [2W][T+1][S][F] (both heads)."""
        assert obs.is_contiguous() and obs.dtype == torch.float32 and obs.shape[0] == 2 * self.W
        head_stride = obs.stride(0)
        step_stride = obs.stride(1) if seq == self.S else self.F
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_window_scatter(_lib.ptr(feats), _lib.ptr(obs), self.W, steps, seq, self.F,
                                                      ctypes.c_int64(head_stride), ctypes.c_int64(step_stride),
                                                      _lib.stream_ptr(self.device)))

    # ------------------------------------------------------------------ public
    @staticmethod
    def _key(rgb, route_fig, measurements):
        return (rgb.data_ptr(), route_fig.data_ptr(), measurements.data_ptr(), tuple(rgb.shape))

    def prefetch(self, rgb, route_fig, measurements):
        """Start copying the NEXT rollout's distinct frames (pinned host tensors, `unique=True` shapes) to the device on
        the copy stream and return immediately. The copy runs under whatever the caller enqueues next (the update phase
        of the current rollout); the following `encode()` of the same tensors consumes the device copies instead of
        staging chunk by chunk. One rollout can be in flight; the copy waits for the previous consumer of the landing
        buffers."""
        if rgb.is_cuda or not (rgb.is_pinned() and route_fig.is_pinned() and measurements.is_pinned()):
            raise _lib.CadreError("prefetch takes frames in pinned host memory")
        if tuple(rgb.shape[:2]) != (self.W, self.K):
            raise _lib.CadreError(f"frames have leading shape {tuple(rgb.shape[:2])}, expected {(self.W, self.K)}")
        n = self.W * self.K
        with torch.cuda.device(self.device):
            if self._landing is None:
                u8 = dict(dtype=torch.uint8, device=self.device)
                self._landing = (torch.empty(n, 144, 256, 3, **u8), torch.empty(n, 256, 144, **u8),
                                 torch.empty(n, 3, dtype=torch.float64, device=self.device))
            self.copy_stream.wait_event(self._landing_free_ev)
            with torch.cuda.stream(self.copy_stream):
                self._landing[0].copy_(rgb.view(n, 144, 256, 3), non_blocking=True)
                self._landing[1].copy_(route_fig.view(n, 256, 144), non_blocking=True)
                self._landing[2].copy_(measurements.view(n, 3), non_blocking=True)
                self._landed_ev.record(self.copy_stream)
        self._landing_key = self._key(rgb, route_fig, measurements)
        self._prefetched_bytes = int(rgb.numel() + route_fig.numel() + measurements.numel() * 8)

    def encode(self, rgb, route_fig, measurements, obs, unique=True):
        """Fill obs[2w + head][t] (t < T) for all workers of this rank from raw frames.

        unique=True : rgb u8 [W, T+S-1, 144, 256, 3], route_fig u8 [W, T+S-1, 256, 144], measurements f64 [W, T+S-1, 3]:
                      every worker's DISTINCT frames, oldest first (tick t sees frames t .. t+S-1);
        unique=False: rgb u8 [W, T, S, 144, 256, 3] (etc.): every tick's full window, as the env wrapper hands it out.
        Inputs may be pinned host tensors (copied H2D here, overlapped with the encoder) or device tensors."""
        host = not rgb.is_cuda
        if host and not (rgb.is_pinned() and route_fig.is_pinned() and measurements.is_pinned()):
            raise _lib.CadreError("host frames must be in pinned memory (torch.empty(..., pin_memory=True))")
        lead = (self.W, self.K) if unique else (self.W, self.T, self.S)
        if tuple(rgb.shape[:len(lead)]) != lead:
            raise _lib.CadreError(f"frames have leading shape {tuple(rgb.shape[:len(lead)])}, expected {lead}")
        n = 1
        for d in lead:
            n *= d
        feats = self._feat_buffer(n)
        if host and unique and self._landing_key is not None and self._landing_key == self._key(rgb, route_fig, measurements):
            # this rollout was prefetched: its frames are (or will be, once the copy stream's event fires) on the device
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._landed_ev)
            self._encode_stream(self._landing[0], self._landing[1], self._landing[2], feats, False)
            self._landing_free_ev.record(main)
            self._landing_key = None
            self._scatter(feats, obs, self.S, self.T)
            self.h2d_bytes_last = self._prefetched_bytes
            return obs
        self._encode_stream(rgb.view(n, 144, 256, 3), route_fig.view(n, 256, 144), measurements.view(n, 3), feats, host)
        if unique:
            self._scatter(feats, obs, self.S, self.T)
        else:
            self._scatter(feats, obs, 1, self.T * self.S)       # rows already are (t, j): a plain 2-head copy
        if not host:
            self.h2d_bytes_last = 0
        return obs
