"""Seeded synthetic weights for runs without a checkpoint (bench.py, smoke, tools, tests): every tensor is generated
from a hash of its parameter NAME, so the same weights can be rebuilt on any box without the reference, and the
reference-generated goldens under tests/golden/ stay valid. Key names follow the reference's
state dicts (carla_perception/Networks/danet.py:86-109, ppo_agent/models.py:101-137, 162-177,
ppo_agent/distributions.py:25-40). No algorithm lives here."""
import hashlib
import math

import torch

FEATURE_DIMS = 530          # config_files/agent_config.py:20
STEER_ACTIONS = 33          # config_files/agent_config.py:1-15
THROTTLE_ACTIONS = 3
HEADS = ("steer", "throttle")


# ----------------------------------------------------------------------------------------------------------
# deterministic fixture weights, keyed by parameter NAME (independent of construction / RNG order, so the
# same tensors can be regenerated on a box that has neither the reference nor the goldens' generator)
def _gen(name, seed):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def fixture_tensor(name, shape, seed, std=1.0, mean=0.0, uniform=None):
    g = _gen(name, seed)
    if uniform is not None:
        lo, hi = uniform
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


_XAVIER = False   # set by danet_fixture_state(init="xavier") while it builds a state dict


def _xavier(name, shape, seed):
    """nn.init.xavier_uniform_: U(-a, a), a = sqrt(6 / (fan_in + fan_out)) (experiments_builder.py:163-188)."""
    rf = 1
    for d in shape[2:]:
        rf *= d
    a = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
    return fixture_tensor(name, shape, seed, uniform=(-a, a))


def _conv_w(sd, name, cout, cin, k, seed, gain=1.0):
    if _XAVIER:
        sd[name] = _xavier(name, (cout, cin, k, k), seed)
        return
    fan_in = cin * k * k
    sd[name] = fixture_tensor(name, (cout, cin, k, k), seed, std=gain * math.sqrt(2.0 / fan_in))


def _bn(sd, prefix, c, seed):
    if _XAVIER:   # experiments_builder.py:177-179 + BatchNorm2d defaults: the identity at eval time (up to eps)
        sd[prefix + ".weight"], sd[prefix + ".bias"] = torch.ones(c), torch.zeros(c)
        sd[prefix + ".running_mean"], sd[prefix + ".running_var"] = torch.zeros(c), torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        return
    # randomised affine + running statistics (defaults would make eval-mode BN ~identity)
    sd[prefix + ".weight"] = fixture_tensor(prefix + ".weight", (c,), seed, uniform=(0.6, 1.4))
    sd[prefix + ".bias"] = fixture_tensor(prefix + ".bias", (c,), seed, std=0.1)
    sd[prefix + ".running_mean"] = fixture_tensor(prefix + ".running_mean", (c,), seed, std=0.1)
    sd[prefix + ".running_var"] = fixture_tensor(prefix + ".running_var", (c,), seed, uniform=(0.6, 1.4))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _linear(sd, prefix, out_f, in_f, seed, gain=1.0, bias_std=0.02):
    if _XAVIER:
        sd[prefix + ".weight"] = _xavier(prefix + ".weight", (out_f, in_f), seed)
        sd[prefix + ".bias"] = torch.zeros(out_f)
        return
    sd[prefix + ".weight"] = fixture_tensor(prefix + ".weight", (out_f, in_f), seed, std=gain / math.sqrt(in_f))
    sd[prefix + ".bias"] = fixture_tensor(prefix + ".bias", (out_f,), seed, std=bias_std)


def danet_fixture_state(seed=0, peaky=False, init="default"):
    """State-dict entries (reference key names) of every DANet tensor on the `get_latent_feature` path.

    `init="xavier"`: the reference trainer's initialisation (Models/experiments_builder.py:163-188): Xavier-uniform
    conv / linear weights, zero biases, identity BatchNorm, and the zero-initialised attention gammas of
    da_att.py:29,61 - i.e. an untrained network, whose activations are orders of magnitude smaller than the default
    fixture's (a dynamic-range test for the fp16 operands).

    Keys follow carla_perception/Networks/danet.py:86-109 (backbone, da_head, visual_conv, bc_conv,
    inter_task_att); decoder heads (visual_branch, bc_branch, in_bc_speed_fc) are off-path and absent.
    `peaky=True` scales the attention projections so the three softmaxes are far from uniform.
    """
    global _XAVIER
    _XAVIER = init == "xavier"
    try:
        return _danet_state(seed, peaky)
    finally:
        _XAVIER = False


def _danet_state(seed, peaky):
    sd = {}
    _conv_w(sd, "backbone.conv1.weight", 64, 4, 7, seed)
    sd["backbone.conv1.bias"] = (torch.zeros(64) if _XAVIER else
                                 fixture_tensor("backbone.conv1.bias", (64,), seed, std=0.05))
    _bn(sd, "backbone.bn1", 64, seed)
    inpl = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}"
            cin = inpl if bi == 0 else planes
            _conv_w(sd, p + ".conv1.weight", planes, cin, 3, seed)
            _bn(sd, p + ".bn1", planes, seed)
            _conv_w(sd, p + ".conv2.weight", planes, planes, 3, seed, gain=0.7)
            _bn(sd, p + ".bn2", planes, seed)
            if bi == 0 and (li > 1):
                _conv_w(sd, p + ".downsample.0.weight", planes, cin, 1, seed, gain=0.7)
                _bn(sd, p + ".downsample.1", planes, seed)
        inpl = planes
    for nm, cin in (("conv5a", 512), ("conv5c", 512), ("conv51", 128), ("conv52", 128)):
        _conv_w(sd, f"da_head.{nm}.0.weight", 128, cin, 3, seed)
        _bn(sd, f"da_head.{nm}.1", 128, seed)
    qk_gain = 1.5 if peaky else 1.0
    for nm, co in (("query_conv", 16), ("key_conv", 16), ("value_conv", 128)):
        g = qk_gain if nm != "value_conv" else 1.0
        if _XAVIER:
            sd[f"da_head.sa.{nm}.weight"] = _xavier(f"da_head.sa.{nm}.weight", (co, 128, 1, 1), seed)
            sd[f"da_head.sa.{nm}.bias"] = torch.zeros(co)
            continue
        sd[f"da_head.sa.{nm}.weight"] = fixture_tensor(f"da_head.sa.{nm}.weight", (co, 128, 1, 1), seed,
                                                        std=g / math.sqrt(128))
        sd[f"da_head.sa.{nm}.bias"] = fixture_tensor(f"da_head.sa.{nm}.bias", (co,), seed, std=0.05)
    sd["da_head.sa.gamma"] = torch.tensor([0.0 if _XAVIER else 0.7])   # zero-init in the reference (da_att.py:29) = no-op
    sd["da_head.sc.gamma"] = torch.tensor([0.0 if _XAVIER else 0.3])   # da_att.py:61
    if _XAVIER:
        sd["da_head.conv8.1.weight"] = _xavier("da_head.conv8.1.weight", (512, 128, 1, 1), seed)
        sd["da_head.conv8.1.bias"] = torch.zeros(512)
    else:
        sd["da_head.conv8.1.weight"] = fixture_tensor("da_head.conv8.1.weight", (512, 128, 1, 1), seed,
                                                      std=1.0 / math.sqrt(128))
        sd["da_head.conv8.1.bias"] = fixture_tensor("da_head.conv8.1.bias", (512,), seed, std=0.05)
    for nm in ("visual_conv", "bc_conv"):
        if _XAVIER:
            sd[nm + ".weight"], sd[nm + ".bias"] = _xavier(nm + ".weight", (512, 512, 1, 1), seed), torch.zeros(512)
            continue
        sd[nm + ".weight"] = fixture_tensor(nm + ".weight", (512, 512, 1, 1), seed, std=1.0 / math.sqrt(512))
        sd[nm + ".bias"] = fixture_tensor(nm + ".bias", (512,), seed, std=0.05)
    it_gain = 2.0 if peaky else 1.0
    for task in ("visual", "bc"):
        for role in ("query", "key", "value"):
            p = f"inter_task_att.{task}_{role}_layer"
            g = it_gain if role != "value" else 1.0
            _linear(sd, p + ".1", 512, 20480, seed)
            _linear(sd, p + ".3", 256, 512, seed, gain=g)
    return sd


PPO_MODULE_ORDER = (
    # dict insertion order of ppo_agent/models.py:101-125 (LSTMs are created inside the first command
    # iteration because the nested loop reuses `_command`)
    ["steer_ppo_0", "throttle_ppo_0"]
    + [f"{h}_lstm_{c}" for c in range(4) for h in HEADS]
    + [f"{h}_ppo_{c}" for c in range(1, 4) for h in HEADS]
)


def ppo_module_param_shapes(name):
    """named_parameters() order and shapes of one PPO module (models.py:130-137 LSTM; :162-177 Model with
    distributions.py:25-40 Categorical_1d registered first as `control`)."""
    F_ = FEATURE_DIMS
    if "_lstm_" in name:
        return [("rnn.weight_ih", (4 * F_, F_)), ("rnn.weight_hh", (4 * F_, F_)),
                ("rnn.bias_ih", (4 * F_,)), ("rnn.bias_hh", (4 * F_,))]
    A = STEER_ACTIONS if name.startswith("steer") else THROTTLE_ACTIONS
    return [("control.linear.0.weight", (128, F_)), ("control.linear.0.bias", (128,)),
            ("control.linear.2.weight", (128, 128)), ("control.linear.2.bias", (128,)),
            ("control.linear.4.weight", (A, 128)), ("control.linear.4.bias", (A,)),
            ("critic.0.weight", (128, F_)), ("critic.0.bias", (128,)),
            ("critic.2.weight", (128, 128)), ("critic.2.bias", (128,)),
            ("critic.4.weight", (1, 128)), ("critic.4.bias", (1,))]


def ppo_fixture_state(seed=0):
    """{module name: {param name: tensor}} for the 16 PPO modules / 128 tensors / 19 382 808 parameters."""
    out = {}
    for m in PPO_MODULE_ORDER:
        sd = {}
        for pn, shape in ppo_module_param_shapes(m):
            full = f"{m}.{pn}"
            if "bias" in pn:
                sd[pn] = fixture_tensor(full, shape, seed, std=0.05)
            else:
                gain = 0.3 if pn.startswith("control.linear.4") else 1.0
                sd[pn] = fixture_tensor(full, shape, seed, std=gain / math.sqrt(shape[1]))
        out[m] = sd
    return out
