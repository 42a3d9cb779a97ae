"""Flat parameter layout of the 16 PPO modules (mirror of cadre_b200/csrc/ppo_layout.h) and conversion from /
to reference-style state dicts (ppo_agent/models.py:44-126: `steer_ppo_c`, `throttle_ppo_c`, `steer_lstm_c`,
`throttle_lstm_c`, c = 0..3; 128 tensors, 19 382 808 parameters).

The same layout is used for parameters, gradients and both Adam moments. Host-side logic only (torch index
arithmetic, runs on any device); the kernels that consume the buffers live in the CUDA library.
"""
import torch

E, F, LDF, G, HID, AMAX, B3A_LD = 8, 530, 532, 2120, 128, 33, 36
HEADS = ("steer", "throttle")
ACTIONS = {"steer": 33, "throttle": 3}

# LSTM tensors are expert-major: block e = [W_ih | W_hh | b_ih | b_hh] of expert e (see ppo_layout.h); the
# actor-critic tensors follow, kind-major.
LSTM_WIH, LSTM_WHH, LSTM_BIH, LSTM_BHH = 0, G * LDF, 2 * G * LDF, 2 * G * LDF + G
LSTM_BLK = 2 * G * LDF + 2 * G
_sizes = [
    ("LSTM", E * LSTM_BLK),
    ("W1", E * 2 * HID * LDF), ("B1", E * 2 * HID), ("W2", E * 2 * HID * HID), ("B2", E * 2 * HID),
    ("W3A", E * AMAX * HID), ("B3A", E * B3A_LD), ("W3C", E * HID), ("B3C", E * 4),
]
OFF = {}
_o = 0
for _n, _s in _sizes:
    OFF[_n] = _o
    _o += _s
TOTAL = _o
NUM_REFERENCE_PARAMS = 19382808

# dict insertion order of the reference's create_model (models.py:101-125)
MODULE_ORDER = (["steer_ppo_0", "throttle_ppo_0"]
                + [f"{h}_lstm_{c}" for c in range(4) for h in HEADS]
                + [f"{h}_ppo_{c}" for c in range(1, 4) for h in HEADS])


def expert_of(module_name):
    head, _, c = module_name.split("_")
    return HEADS.index(head) * 4 + int(c)


def module_param_names(module_name):
    """named_parameters() order of a reference module (LSTM: models.py:130-137; Model: models.py:162-177 with
    distributions.py:25-40 `control` registered before `critic`)."""
    if "_lstm_" in module_name:
        return ["rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"]
    return ["control.linear.0.weight", "control.linear.0.bias", "control.linear.2.weight", "control.linear.2.bias",
            "control.linear.4.weight", "control.linear.4.bias", "critic.0.weight", "critic.0.bias",
            "critic.2.weight", "critic.2.bias", "critic.4.weight", "critic.4.bias"]


def _gate_interleave(t):
    """[4*F, ...] gate-major (i,f,g,o blocks, torch LSTMCell) -> row 4*u + g."""
    return t.reshape(4, F, *t.shape[1:]).transpose(0, 1).reshape(4 * F, *t.shape[1:])


def _gate_deinterleave(t):
    return t.reshape(F, 4, *t.shape[1:]).transpose(0, 1).reshape(4 * F, *t.shape[1:])


def tensor_view(flat, module_name, param_name):
    """(view into `flat`, needs_gate_permutation) for one reference tensor. The view has the reference shape
    except for LSTM tensors, whose rows are gate-interleaved (convert with _gate_(de)interleave)."""
    e = expert_of(module_name)
    if "_lstm_" in module_name:
        base = OFF["LSTM"] + e * LSTM_BLK
        if param_name in ("rnn.weight_ih", "rnn.weight_hh"):
            base += LSTM_WIH if param_name == "rnn.weight_ih" else LSTM_WHH
            v = flat[base: base + G * LDF].view(G, LDF)[:, :F]
        else:
            base += LSTM_BIH if param_name == "rnn.bias_ih" else LSTM_BHH
            v = flat[base: base + G]
        return v, True
    A = ACTIONS[module_name.split("_")[0]]
    br = 0 if param_name.startswith("control") else 1
    layer = param_name.split(".")[-2]
    is_w = param_name.endswith("weight")
    if layer == "0":
        if is_w:
            base = OFF["W1"] + (e * 2 + br) * HID * LDF
            return flat[base: base + HID * LDF].view(HID, LDF)[:, :F], False
        base = OFF["B1"] + (e * 2 + br) * HID
        return flat[base: base + HID], False
    if layer == "2":
        if is_w:
            base = OFF["W2"] + (e * 2 + br) * HID * HID
            return flat[base: base + HID * HID].view(HID, HID), False
        base = OFF["B2"] + (e * 2 + br) * HID
        return flat[base: base + HID], False
    if br == 0:
        if is_w:
            base = OFF["W3A"] + e * AMAX * HID
            return flat[base: base + A * HID].view(A, HID), False
        base = OFF["B3A"] + e * B3A_LD
        return flat[base: base + A], False
    if is_w:
        base = OFF["W3C"] + e * HID
        return flat[base: base + HID].view(1, HID), False
    base = OFF["B3C"] + e * 4
    return flat[base: base + 1], False


def pack_state(state, device="cpu"):
    """{module: {param: tensor}} (reference shapes) -> flat fp32 buffer [TOTAL]."""
    flat = torch.zeros(TOTAL, dtype=torch.float32, device=device)
    for m in MODULE_ORDER:
        for pn in module_param_names(m):
            view, perm = tensor_view(flat, m, pn)
            t = state[m][pn].detach().to(device=device, dtype=torch.float32)
            view.copy_(_gate_interleave(t) if perm else t)
    return flat


def unpack_state(flat):
    """flat buffer -> {module: {param: tensor}} with reference shapes (copies)."""
    out = {}
    for m in MODULE_ORDER:
        d = {}
        for pn in module_param_names(m):
            view, perm = tensor_view(flat, m, pn)
            d[pn] = (_gate_deinterleave(view) if perm else view).clone()
        out[m] = d
    return out
