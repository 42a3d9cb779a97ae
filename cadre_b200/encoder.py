"""Host side of the perception encoder: one-time weight preparation and the ctypes call into
libcadre_sm100.so. Mirrors `DANet.get_latent_feature(x, "concate")` (carla_perception/Networks/danet.py:216-238)
in eval mode as loaded by `create_model(load_vae=True)` (ppo_agent/models.py:54-88).

All compute happens in the CUDA library; torch is used for device memory and for the one-time (load-time)
algebra on the weights:
  * BatchNorm (eval) folded into the preceding convolution (resnet.py:43-50, danet.py:21-36);
  * conv weights permuted to [Cout][KH][KW][Cin] (NHWC implicit GEMM), bf16;
  * the linear chain conv8 -> visual_conv / bc_conv -> Linear(20480, 512) (danet.py:41,222,226;
    intertask_att.py:39-80; Dropout2d is the identity in eval mode) folded into one [3072, 5120] matrix.
"""
import ctypes

import torch

from . import _lib

_PTR = ctypes.c_void_p


class EncoderWeights(ctypes.Structure):
    """Mirror of `cadre_encoder_weights` (include/cadre_b200.h)."""
    _fields_ = [
        ("stem_w", _PTR), ("stem_b", _PTR),
        ("conv_w", _PTR * 19), ("conv_b", _PTR * 19),
        ("head5_w", _PTR), ("head5_b", _PTR),
        ("pam_wqk", _PTR), ("pam_bqk", _PTR), ("pam_wv", _PTR), ("pam_bv", _PTR),
        ("conv51_w", _PTR), ("conv51_b", _PTR), ("conv52_w", _PTR), ("conv52_b", _PTR),
        ("fc1_w", _PTR), ("fc1_b", _PTR), ("fc2_w", _PTR), ("fc2_b", _PTR),
        ("pam_gamma", ctypes.c_float), ("cam_gamma", ctypes.c_float),
    ]


def _fold_bn(w, b, sd, bn, eps=1e-5):
    """conv weight [Cout,Cin,KH,KW] (+ optional bias) followed by eval-mode BatchNorm2d -> (w', b') in fp64."""
    w = w.double()
    scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + eps)
    b0 = b.double() if b is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    return w * scale[:, None, None, None], (b0 - sd[bn + ".running_mean"].double()) * scale + sd[bn + ".bias"].double()


def _khwc(w):
    return w.permute(0, 2, 3, 1).contiguous().reshape(w.shape[0], -1)


# fp16 operands: largest finite value and smallest NORMAL magnitude (below it precision degrades bit by bit)
FP16_MAX, FP16_MIN_NORMAL = 65504.0, 6.103515625e-05
SUBNORMAL_FRACTION_LIMIT = 0.01


def required_keys():
    """Names of the checkpoint tensors the `get_latent_feature` path reads (danet.py:86-109, 216-238)."""
    keys = ["backbone.conv1.weight", "backbone.conv1.bias"]
    bn = lambda p: [p + s for s in (".weight", ".bias", ".running_mean", ".running_var")]  # noqa: E731
    keys += bn("backbone.bn1")
    for li in range(1, 5):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}"
            keys += [p + ".conv1.weight", p + ".conv2.weight"] + bn(p + ".bn1") + bn(p + ".bn2")
            if li > 1 and bi == 0:
                keys += [p + ".downsample.0.weight"] + bn(p + ".downsample.1")
    for nm in ("conv5a", "conv5c", "conv51", "conv52"):
        keys += [f"da_head.{nm}.0.weight"] + bn(f"da_head.{nm}.1")
    for nm in ("query_conv", "key_conv", "value_conv"):
        keys += [f"da_head.sa.{nm}.weight", f"da_head.sa.{nm}.bias"]
    keys += ["da_head.sa.gamma", "da_head.sc.gamma", "da_head.conv8.1.weight", "da_head.conv8.1.bias",
             "visual_conv.weight", "visual_conv.bias", "bc_conv.weight", "bc_conv.bias"]
    for task in ("visual", "bc"):
        for role in ("query", "key", "value"):
            p = f"inter_task_att.{task}_{role}_layer"
            keys += [p + ".1.weight", p + ".1.bias", p + ".3.weight", p + ".3.bias"]
    return keys


def prepare_weights(sd, device, enc16=torch.float16, check_range=True):
    """DANet state dict (reference key names) -> dict of device tensors in the library's layouts.
    `enc16` is the library's 16-bit operand type (`_lib.enc_dtype()`).

    The folds run in fp64 and the results are rounded ONCE to fp16, without per-channel scales. That is safe as long as
    a folded matrix fits the fp16 range, which `check_range` verifies instead of assuming: a folded weight beyond
    65504 (a BatchNorm with a vanishing running_var, say) or a matrix with more than 1 % of its non-zero entries below
    the smallest normal fp16 raises CadreError naming the tensor - never a silently saturated / flushed operand."""
    missing = [k for k in required_keys() if k not in sd]
    if missing:
        raise _lib.CadreError(f"perception checkpoint lacks {len(missing)} tensor(s) of the get_latent_feature path, "
                              f"e.g. {missing[:4]}")
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    out = {}

    def put(name, t, dtype):
        if check_range and dtype == torch.float16:
            a = t.detach().abs().double()
            if not torch.isfinite(a).all() or float(a.max()) > FP16_MAX:
                raise _lib.CadreError(f"folded weight `{name}` does not fit fp16 (max |w| = {float(a.max()):.3e} > "
                                      f"{FP16_MAX}); the checkpoint's BatchNorm statistics drive it out of range")
            nz = a > 0
            sub = float(((a < FP16_MIN_NORMAL) & nz).sum()) / max(1, int(nz.sum()))
            if sub > SUBNORMAL_FRACTION_LIMIT:
                raise _lib.CadreError(f"folded weight `{name}`: {100 * sub:.1f} % of its non-zero entries are below the "
                                      f"smallest normal fp16 ({FP16_MIN_NORMAL:.2e}) and would lose precision")
        out[name] = t.to(dtype).contiguous().to(device)

    # stem (resnet.py:111-112,169-171): conv7x7 s2 with bias + BN
    w, b = _fold_bn(sd["backbone.conv1.weight"], sd["backbone.conv1.bias"], sd, "backbone.bn1")
    wk = torch.zeros(64, 8, 8, 4, dtype=torch.float64)
    wk[:, :7, :7, :] = w.permute(0, 2, 3, 1)
    wk = wk.view(64, 4, 2, 8, 4).permute(0, 1, 3, 2, 4).contiguous().view(64, 256)
    put("stem_w", wk, enc16)
    put("stem_b", b, torch.float32)

    convs = []
    for li in range(1, 5):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}"
            convs.append(_fold_bn(sd[p + ".conv1.weight"], None, sd, p + ".bn1"))
            convs.append(_fold_bn(sd[p + ".conv2.weight"], None, sd, p + ".bn2"))
            if (p + ".downsample.0.weight") in sd:
                convs.append(_fold_bn(sd[p + ".downsample.0.weight"], None, sd, p + ".downsample.1"))
    assert len(convs) == 19
    for i, (w, b) in enumerate(convs):
        put(f"conv_w{i}", _khwc(w), enc16)
        put(f"conv_b{i}", b, torch.float32)

    wa, ba = _fold_bn(sd["da_head.conv5a.0.weight"], None, sd, "da_head.conv5a.1")
    wc, bc = _fold_bn(sd["da_head.conv5c.0.weight"], None, sd, "da_head.conv5c.1")
    put("head5_w", torch.cat([_khwc(wa), _khwc(wc)], 0), enc16)
    put("head5_b", torch.cat([ba, bc], 0), torch.float32)
    put("pam_wqk", torch.cat([sd["da_head.sa.query_conv.weight"].view(16, 128),
                              sd["da_head.sa.key_conv.weight"].view(16, 128)], 0), torch.float32)
    put("pam_bqk", torch.cat([sd["da_head.sa.query_conv.bias"], sd["da_head.sa.key_conv.bias"]], 0), torch.float32)
    put("pam_wv", sd["da_head.sa.value_conv.weight"].view(128, 128), enc16)
    put("pam_bv", sd["da_head.sa.value_conv.bias"], torch.float32)
    for nm in ("conv51", "conv52"):
        w, b = _fold_bn(sd[f"da_head.{nm}.0.weight"], None, sd, f"da_head.{nm}.1")
        put(nm + "_w", _khwc(w), enc16)
        put(nm + "_b", b, torch.float32)

    # fold conv8 -> task conv -> Linear1 (all linear; flatten order of the reference is NCHW: c*40 + h*8 + w)
    w8 = sd["da_head.conv8.1.weight"].view(512, 128).double()
    b8 = sd["da_head.conv8.1.bias"].double()
    fc1_w, fc1_b, fc2_w, fc2_b = [], [], [], []
    for task, conv in (("visual", "visual_conv"), ("bc", "bc_conv")):
        wt = sd[conv + ".weight"].view(512, 512).double()
        m = wt @ w8                                   # [c', c128]
        bias_c = wt @ b8 + sd[conv + ".bias"].double()  # [c']
        for role in ("query", "key", "value"):
            p = f"inter_task_att.{task}_{role}_layer"
            w1 = sd[p + ".1.weight"].double().view(512, 512, 40)     # [o, c', hw]
            fc1_w.append(torch.einsum("ock,cd->okd", w1, m).reshape(512, 5120))
            fc1_b.append(sd[p + ".1.bias"].double() + torch.einsum("ock,c->o", w1, bias_c))
            fc2_w.append(sd[p + ".3.weight"].double())
            fc2_b.append(sd[p + ".3.bias"].double())
    put("fc1_w", torch.cat(fc1_w, 0), enc16)
    put("fc1_b", torch.cat(fc1_b, 0), torch.float32)
    put("fc2_w", torch.stack(fc2_w, 0), enc16)
    put("fc2_b", torch.stack(fc2_b, 0), torch.float32)
    out["pam_gamma"] = float(sd["da_head.sa.gamma"].item())
    out["cam_gamma"] = float(sd["da_head.sc.gamma"].item())
    return out


class Encoder:
    """B200 perception encoder. `state_dict` uses the reference DANet key names (the `'autoencoder'` entry of
    the reference checkpoint, models.py:55-63)."""

    def __init__(self, state_dict, device="cuda:0", max_batch=1024):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CadreError("cadre_b200.Encoder needs a CUDA device: there is no CPU fallback")
        self.max_batch = int(max_batch)
        self._lib = _lib.lib()
        with torch.cuda.device(self.device):
            self.enc16 = _lib.enc_dtype()
            self._t = prepare_weights(state_dict, self.device, self.enc16)
            w = EncoderWeights()
            for name, _ in EncoderWeights._fields_:
                if name in ("conv_w", "conv_b"):
                    arr = getattr(w, name)
                    for i in range(19):
                        arr[i] = self._t[f"{name}{i}"].data_ptr()
                elif name in ("pam_gamma", "cam_gamma"):
                    setattr(w, name, self._t[name])
                else:
                    setattr(w, name, self._t[name].data_ptr())
            self._w = w
            h = ctypes.c_void_p()
            _lib.check(self._lib.cadre_encoder_create(ctypes.byref(h), ctypes.byref(w), self.max_batch))
            self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                with torch.cuda.device(self.device):
                    self._lib.cadre_encoder_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward_u8(self, rgb, route_fig, measurements=None, out=None):
        """rgb u8 [B,144,256,3], route_fig u8 [B,256,144], measurements f64 [B,3] (all on the device).
        Returns fp32 [B,530] (latent + 18 measurement floats) or [B,512] without measurements."""
        B = rgb.shape[0]
        width = 530 if measurements is not None else 512
        if out is None:
            out = torch.empty(B, width, device=self.device, dtype=torch.float32)
        assert rgb.dtype == torch.uint8 and route_fig.dtype == torch.uint8 and rgb.is_contiguous()
        assert route_fig.is_contiguous() and out.stride(1) == 1
        if measurements is not None:
            assert measurements.dtype == torch.float64 and measurements.is_contiguous()
        with torch.cuda.device(self.device):     # the library launches on the CURRENT device (vae_device may differ)
            for s in range(0, B, self.max_batch):
                n = min(self.max_batch, B - s)
                _lib.check(self._lib.cadre_encoder_forward_u8(
                    self._h, _lib.ptr(rgb[s:]), _lib.ptr(route_fig[s:]),
                    _lib.ptr(measurements[s:]) if measurements is not None else None, n, _lib.ptr(out[s:]),
                    out.stride(0), _lib.stream_ptr(self.device)))
        return out

    def forward_f32(self, x, out=None):
        """x fp32 NCHW [B,4,144,256] already pre-processed. Returns fp32 [B,512]."""
        B = x.shape[0]
        assert x.dtype == torch.float32 and tuple(x.shape[1:]) == (4, 144, 256)
        x = x.contiguous()
        if out is None:
            out = torch.empty(B, 512, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            for s in range(0, B, self.max_batch):
                n = min(self.max_batch, B - s)
                _lib.check(self._lib.cadre_encoder_forward_f32(self._h, _lib.ptr(x[s:]), n, _lib.ptr(out[s:]),
                                                               out.stride(0), _lib.stream_ptr(self.device)))
        return out

    def debug_buffer(self, which, B):
        """View of an internal activation buffer (see cadre_encoder_buffer) for the first B frames."""
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        _lib.check(self._lib.cadre_encoder_buffer(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        t = torch.empty(B * n.value, device=self.device, dtype=self.enc16)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_memcpy_d2d(_lib.ptr(t), p, ctypes.c_int64(B * n.value * 2),
                                                  _lib.stream_ptr(self.device)))
        return t

    def activation_absmax(self, B):
        """max |activation| of the first B frames in the debug-visible buffers of the last forward (layer4, conv5a|5c,
        PAM, CAM, head sum, fc1). 16-bit stores SATURATE at 65504 (fp16), so a value at the limit means the checkpoint
        drives activations out of the fp16 range."""
        return max(float(self.debug_buffer(w, B).float().abs().max()) for w in range(6))

    def profile(self, B):
        """[(launch name, ms)] for the trunk re-run on the B frames ingested by the previous forward call."""
        out = torch.empty(B, 512, device=self.device, dtype=torch.float32)
        ms = (ctypes.c_float * 64)()
        names = ctypes.create_string_buffer(4096)
        n = ctypes.c_int()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_encoder_profile(self._h, B, _lib.ptr(out), 512, ms, names, 4096,
                                                       ctypes.byref(n), _lib.stream_ptr(self.device)))
        return list(zip(names.value.decode().split(";"), [ms[i] for i in range(n.value)]))

    @property
    def launches_per_forward(self):
        return int(self._lib.cadre_encoder_launches(self._h))
