"""ORACLE — TEST INFRASTRUCTURE ONLY. Not shipped, not imported by the product package `cadre_b200`.

CPU restatement (plain PyTorch fp32, functional style over state dicts) of the CADRE learner hot path of
BIT-MCS/Cadre, function by function. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module, and only as the checker / reported CPU baseline.

Pinning status: the reference ships NO tests, golden vectors or seeds for this path, so parity
is pinned by running the reference's own modules in the build container: `oracle/ref_shim.py` imports them
from /root/reference, `oracle/make_golden.py` writes `tests/golden/*.npz` from those reference runs, and
`tests/test_oracle.py` checks this restatement against (i) the committed goldens everywhere and (ii) the live
reference, bit for bit, whenever /root/reference is present.

Every function cites the reference file:line it restates (paths relative to the reference root).
"""
import hashlib
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------------
# sizes (config_files/agent_config.py:1-58, carla_perception/Config/auto_danet.py:111-149)
NUM_STEPS = 200
MINI_BATCH_NUM = 2
FEATURE_DIMS = 530
SEQ_LENGTH = 8
GAMMA = 0.99
TAU = 0.95
COMMAND_NUM = 4
STEER_ACTIONS = 33
THROTTLE_ACTIONS = 3
ENT_COEFF, VALUE_COEFF, CLIP_COEFF, CLIP = 0.01, 0.1, 1.0, 0.1
LR, MAX_GRAD_NORM, PPO_EPOCH = 3e-4, 250.0, 4
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


def _conv_w(sd, name, cout, cin, k, seed, gain=1.0):
    fan_in = cin * k * k
    sd[name] = fixture_tensor(name, (cout, cin, k, k), seed, std=gain * math.sqrt(2.0 / fan_in))


def _bn(sd, prefix, c, seed):
    # randomised affine + running statistics (defaults would make eval-mode BN ~identity)
    sd[prefix + ".weight"] = fixture_tensor(prefix + ".weight", (c,), seed, uniform=(0.6, 1.4))
    sd[prefix + ".bias"] = fixture_tensor(prefix + ".bias", (c,), seed, std=0.1)
    sd[prefix + ".running_mean"] = fixture_tensor(prefix + ".running_mean", (c,), seed, std=0.1)
    sd[prefix + ".running_var"] = fixture_tensor(prefix + ".running_var", (c,), seed, uniform=(0.6, 1.4))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _linear(sd, prefix, out_f, in_f, seed, gain=1.0, bias_std=0.02):
    sd[prefix + ".weight"] = fixture_tensor(prefix + ".weight", (out_f, in_f), seed, std=gain / math.sqrt(in_f))
    sd[prefix + ".bias"] = fixture_tensor(prefix + ".bias", (out_f,), seed, std=bias_std)


def danet_fixture_state(seed=0, peaky=False):
    """State-dict entries (reference key names) of every DANet tensor on the `get_latent_feature` path.

    Keys follow carla_perception/Networks/danet.py:86-109 (backbone, da_head, visual_conv, bc_conv,
    inter_task_att); decoder heads (visual_branch, bc_branch, in_bc_speed_fc) are off-path and absent.
    `peaky=True` scales the attention projections so the three softmaxes are far from uniform.
    """
    sd = {}
    _conv_w(sd, "backbone.conv1.weight", 64, 4, 7, seed)
    sd["backbone.conv1.bias"] = fixture_tensor("backbone.conv1.bias", (64,), seed, std=0.05)
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
        sd[f"da_head.sa.{nm}.weight"] = fixture_tensor(f"da_head.sa.{nm}.weight", (co, 128, 1, 1), seed,
                                                        std=g / math.sqrt(128))
        sd[f"da_head.sa.{nm}.bias"] = fixture_tensor(f"da_head.sa.{nm}.bias", (co,), seed, std=0.05)
    sd["da_head.sa.gamma"] = torch.tensor([0.7])   # zero-init in the reference (da_att.py:29) = no-op
    sd["da_head.sc.gamma"] = torch.tensor([0.3])   # da_att.py:61
    sd["da_head.conv8.1.weight"] = fixture_tensor("da_head.conv8.1.weight", (512, 128, 1, 1), seed,
                                                  std=1.0 / math.sqrt(128))
    sd["da_head.conv8.1.bias"] = fixture_tensor("da_head.conv8.1.bias", (512,), seed, std=0.05)
    for nm in ("visual_conv", "bc_conv"):
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


# ----------------------------------------------------------------------------------------------------------
# encoder
def pre_process(rgb_u8, route_u8):
    """ppo_agent/agent.py:43-75 (use_vae branch). rgb u8 [S,144,256,3], route_fig u8 [S,256,144].

    The route map is max-normalised per frame and WRITTEN BACK INTO THE uint8 ARRAY (agent.py:51-54), i.e.
    truncated to {0,1}; like the reference this mutates `route_u8` in place.
    """
    rgb = np.array(rgb_u8 / 255., dtype=np.float32)
    img = rgb.transpose(0, 3, 1, 2)
    for i in range(route_u8.shape[0]):
        mx = np.max(route_u8[i]) * 1.0
        if mx > 0:
            route_u8[i] = 1.0 * route_u8[i] / mx
    route = np.array(route_u8, dtype=np.float32).swapaxes(1, 2)
    route = np.expand_dims(route, 1)
    return np.concatenate([img, route], axis=1)


def _bn_eval(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def _basic_block(x, sd, p, stride):
    """danet_blocks/resnet.py:39-55."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    out = F.relu(_bn_eval(out, sd, p + ".bn1"))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn_eval(out, sd, p + ".bn2")
    idn = x
    if (p + ".downsample.0.weight") in sd:
        idn = _bn_eval(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    out = out + idn
    return F.relu(out)


def backbone(x, sd):
    """danet_blocks/resnet.py:168-183 (ResNet-18, bias_first stem)."""
    x = F.conv2d(x, sd["backbone.conv1.weight"], sd["backbone.conv1.bias"], 2, 3)
    x = F.relu(_bn_eval(x, sd, "backbone.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        x = _basic_block(x, sd, f"backbone.layer{li}.0", 1 if li == 1 else 2)
        x = _basic_block(x, sd, f"backbone.layer{li}.1", 1)
    return x


def pam(x, sd, p="da_head.sa"):
    """danet_blocks/da_att.py:32-51."""
    b, c, h, w = x.shape
    q = F.conv2d(x, sd[p + ".query_conv.weight"], sd[p + ".query_conv.bias"]).view(b, -1, h * w).permute(0, 2, 1)
    k = F.conv2d(x, sd[p + ".key_conv.weight"], sd[p + ".key_conv.bias"]).view(b, -1, h * w)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(x, sd[p + ".value_conv.weight"], sd[p + ".value_conv.bias"]).view(b, -1, h * w)
    out = torch.bmm(v, att.permute(0, 2, 1)).view(b, c, h, w)
    return sd[p + ".gamma"] * out + x


def cam(x, sd, p="da_head.sc"):
    """danet_blocks/da_att.py:63-83."""
    b, c, h, w = x.shape
    xq = x.view(b, c, -1)
    energy = torch.bmm(xq, xq.permute(0, 2, 1))
    energy_new = torch.max(energy, -1, keepdim=True)[0].expand_as(energy) - energy
    att = torch.softmax(energy_new, dim=-1)
    out = torch.bmm(att, xq).view(b, c, h, w)
    return sd[p + ".gamma"] * out + x


def _conv_bn_relu(x, sd, p):
    return F.relu(_bn_eval(F.conv2d(x, sd[p + ".0.weight"], None, 1, 1), sd, p + ".1"))


def da_head(x, sd):
    """danet.py:43-69 (Dropout2d of conv8 is the identity in eval mode, models.py:86)."""
    sa = _conv_bn_relu(pam(_conv_bn_relu(x, sd, "da_head.conv5a"), sd), sd, "da_head.conv51")
    sc = _conv_bn_relu(cam(_conv_bn_relu(x, sd, "da_head.conv5c"), sd), sd, "da_head.conv52")
    return F.conv2d(sa + sc, sd["da_head.conv8.1.weight"], sd["da_head.conv8.1.bias"])


def _mlp(x, sd, p):
    """intertask_att.py:39-80: Flatten -> Linear(20480,512) -> LeakyReLU -> Linear(512,256)."""
    h = F.leaky_relu(F.linear(x, sd[p + ".1.weight"], sd[p + ".1.bias"]), 0.01)
    return F.linear(h, sd[p + ".3.weight"], sd[p + ".3.bias"])


def inter_task_att(vis, bc, sd, z_dims=256):
    """intertask_att.py:123-176 (transformer branch; nn.Dropout is the identity in eval mode)."""
    b = vis.shape[0]
    vis, bc = vis.reshape(b, -1), bc.reshape(b, -1)
    p = "inter_task_att."
    vq, vk, vv = (_mlp(vis, sd, p + f"visual_{r}_layer") for r in ("query", "key", "value"))
    bq, bk, bv = (_mlp(bc, sd, p + f"bc_{r}_layer") for r in ("query", "key", "value"))
    temp = z_dims ** 0.5
    e = torch.bmm(vq.view(b, 1, z_dims).permute(0, 2, 1) / temp, bk.view(b, 1, z_dims))
    att = torch.softmax(e, dim=-1)
    att_bc = torch.bmm(bv.view(b, 1, z_dims), att.permute(0, 2, 1)).view(b, -1) + bv
    e = torch.bmm(bq.view(b, 1, z_dims).permute(0, 2, 1) / temp, vk.view(b, 1, z_dims))
    att = torch.softmax(e, dim=-1)
    att_vis = torch.bmm(vv.view(b, 1, z_dims), att.permute(0, 2, 1)).view(b, -1) + vv
    return att_vis, att_bc


def encoder_latent(x, sd):
    """danet.py:216-238 `get_latent_feature(x, "concate")`: [B,4,144,256] fp32 -> [B,512]."""
    l4 = backbone(x, sd)
    da = da_head(l4, sd)
    vis = F.conv2d(da, sd["visual_conv.weight"], sd["visual_conv.bias"])
    bc = F.conv2d(da, sd["bc_conv.weight"], sd["bc_conv.bias"])
    av, ab = inter_task_att(vis, bc, sd)
    return torch.cat((av, ab), dim=-1)


def agent_latent_feature(rgb_u8, route_u8, measurements, sd):
    """ppo_agent/agent.py:97-112: pre_process -> encoder -> cat 18 measurement floats -> float32 [S,530]."""
    x = torch.from_numpy(pre_process(rgb_u8, route_u8))
    lat = encoder_latent(x, sd).clone().detach()
    m = torch.from_numpy(measurements).repeat(1, 6)
    return torch.cat([lat, m], dim=-1).float()


# ----------------------------------------------------------------------------------------------------------
# PPO networks
def lstm_forward(x, h0, c0, sd):
    """ppo_agent/models.py:139-152 (nn.LSTMCell, gate order i,f,g,o); x [T*N,F] time-major or [N,F]."""
    w_ih, w_hh, b_ih, b_hh = sd["rnn.weight_ih"], sd["rnn.weight_hh"], sd["rnn.bias_ih"], sd["rnn.bias_hh"]
    h, c = h0, c0
    if x.size(0) == h0.size(0):
        h, c = torch.lstm_cell(x, (h, c), w_ih, w_hh, b_ih, b_hh)
    else:
        n = h0.size(0)
        t = int(x.size(0) / n)
        xs = x.view(t, n, x.size(1))
        for i in range(t):
            h, c = torch.lstm_cell(xs[i], (h, c), w_ih, w_hh, b_ih, b_hh)
    return h, (h, c)


def _mlp3(x, sd, p):
    h = F.relu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    h = F.relu(F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
    return F.linear(h, sd[p + "4.weight"], sd[p + "4.bias"])


def evaluate_actions(feat, action, sd):
    """models.py:203-212 + distributions.py:66-83,101-105 (torch.distributions.Categorical(logits=...))."""
    value = _mlp3(feat, sd, "critic.")
    logits = _mlp3(feat, sd, "control.linear.")
    logits = logits - logits.logsumexp(dim=-1, keepdim=True)
    probs = F.softmax(logits, dim=-1)
    logp = logits.gather(-1, action.squeeze(-1).long().unsqueeze(-1))
    min_real = torch.finfo(logits.dtype).min
    ent = -(torch.clamp(logits, min=min_real) * probs).sum(-1).unsqueeze(-1)
    return value, logp, ent


def _head_loss(samples, head, params, clip=CLIP):
    """One head's half of ppo_agent/agent.py:166-229: dense loop over the 4 commands, masked by command."""
    obs, action, old_v, ret, _masks, old_lp, adv, (hn, cn), command = samples
    cur_v = cur_lp = ent = 0
    for c in range(COMMAND_NUM):
        feat, _ = lstm_forward(obs.clone(), hn, cn, params[f"{head}_lstm_{c}"])
        sel = command == c
        v, lp, e = evaluate_actions(feat, action, params[f"{head}_ppo_{c}"])
        cur_v = cur_v + v * sel
        cur_lp = cur_lp + lp * sel
        ent = ent + e * sel
    ratio = torch.exp(cur_lp - old_lp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv
    action_loss = -torch.min(surr1, surr2).mean()
    v_clipped = old_v + (cur_v - old_v).clamp(-clip, clip)
    value_loss = 0.5 * torch.max((cur_v - ret).pow(2), (v_clipped - ret).pow(2)).mean()
    return value_loss, action_loss, ent.mean()


def update_policy(steer_samples, throttle_samples, params):
    """ppo_agent/agent.py:166-237. `params` = {module: {name: leaf tensor with requires_grad}}; gradients are
    left in `.grad` of every leaf (zeroed first, like zero_grad + backward). Returns the three Python floats."""
    vs, as_, es = _head_loss(steer_samples, "steer", params)
    vt, at, et = _head_loss(throttle_samples, "throttle", params)
    value_loss = (vs + vt) * VALUE_COEFF
    action_loss = (as_ + at) * CLIP_COEFF
    ent_loss = (es + et) * ENT_COEFF
    total = value_loss + action_loss - ent_loss
    for m in params.values():
        for p in m.values():
            p.grad = None
    total.backward()
    for m in params.values():
        for p in m.values():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
    return value_loss.item(), action_loss.item(), ent_loss.item()


# ----------------------------------------------------------------------------------------------------------
# rollout storage
def compute_returns(rewards, value_preds, masks, next_value, gamma=GAMMA, tau=TAU):
    """ppo_agent/storage.py:68-76 (use_gae branch). All inputs [T+1,1]; returns a new `returns` [T+1,1] and
    the updated value_preds (slot T overwritten with next_value)."""
    T = rewards.shape[0] - 1
    value_preds = value_preds.clone()
    value_preds[-1] = next_value
    returns = torch.zeros_like(rewards)
    gae = 0
    for step in reversed(range(T)):
        delta = rewards[step] + gamma * value_preds[step + 1] * masks[step] - value_preds[step]
        gae = delta + gamma * tau * masks[step] * gae
        returns[step] = gae + value_preds[step]
    return returns, value_preds


def normalized_advantages(returns, value_preds):
    """ppo_agent/train.py:82-88 (unbiased std over the T values of one storage)."""
    adv = returns[:-1] - value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-8)


def minibatch_indices(num_steps=NUM_STEPS, mini_batch_num=MINI_BATCH_NUM):
    """ppo_agent/storage.py:93-97: BatchSampler(SubsetRandomSampler(range(T)), T // mini_batch_num,
    drop_last=False) == one torch.randperm(T) from the global CPU generator, chunked."""
    mb = num_steps // mini_batch_num
    perm = torch.randperm(num_steps).tolist()
    return [perm[i:i + mb] for i in range(0, num_steps, mb)]


def gather_minibatch(st, adv, idx):
    """ppo_agent/storage.py:98-120. `st` is a dict of the RolloutStorage tensors."""
    obs = st["obs"][idx].permute(1, 0, 2)
    obs = obs.reshape(-1, obs.size(-1))
    return (obs, st["action"][idx], st["value_preds"][idx], st["returns"][idx], st["masks"][idx],
            st["action_log_probs"][idx], adv[idx], [st["hn"][idx], st["cn"][idx]], st["command"][idx])


# ----------------------------------------------------------------------------------------------------------
# chief
def chief_step(params, summed_grads, adam_state, step, lr=LR, max_norm=MAX_GRAD_NORM, betas=(0.9, 0.999),
               eps=1e-8):
    """ppo_agent/chief.py:13-21 + torch.optim.Adam (main.py:55): install SUMMED worker grads, clip each
    module's grad norm to 250 separately (clip_coef = max_norm / (norm + 1e-6), clamped to 1), one Adam step
    with bias correction. Updates params / adam_state in place. `step` is the 1-based Adam step."""
    with torch.no_grad():
        for m in PPO_MODULE_ORDER:
            gs = [summed_grads[m][n] for n, _ in ppo_module_param_shapes(m)]
            total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in gs]), 2.0)
            coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
            for n, _ in ppo_module_param_shapes(m):
                g = summed_grads[m][n] * coef
                p = params[m][n]
                st = adam_state[m][n]
                st["exp_avg"].lerp_(g, 1 - betas[0])
                st["exp_avg_sq"].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
                bc1 = 1 - betas[0] ** step
                bc2 = 1 - betas[1] ** step
                denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
                p.addcdiv_(st["exp_avg"], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------------------------------------
# synthetic inputs shared by fixtures
def synthetic_tick(rs, seq=SEQ_LENGTH):
    rgb = rs.randint(0, 256, size=(seq, 144, 256, 3)).astype(np.uint8)
    route = (rs.rand(seq, 256, 144) < 0.1).astype(np.uint8) * 255
    meas = rs.rand(seq, 3)
    return {"rgb": rgb, "route_fig": route, "measurements": meas, "command": int(rs.randint(0, 4))}


def synthetic_storage(rs, T=NUM_STEPS, actions=STEER_ACTIONS, feature_dims=FEATURE_DIMS, seq=SEQ_LENGTH):
    """One pre-filled RolloutStorage as a dict of tensors with the layout of storage.py:8-26."""
    st = {
        "obs": torch.from_numpy(rs.randn(T + 1, seq, feature_dims).astype(np.float32)),
        "rewards": torch.from_numpy(rs.rand(T + 1, 1).astype(np.float32)),
        "value_preds": torch.from_numpy(rs.randn(T + 1, 1).astype(np.float32)),
        "returns": torch.zeros(T + 1, 1),
        "action_log_probs": torch.from_numpy(rs.uniform(-3.0, -0.5, size=(T + 1, 1)).astype(np.float32)),
        "action": torch.from_numpy(rs.randint(0, actions, size=(T + 1, 1)).astype(np.int64)),
        "masks": torch.from_numpy((rs.rand(T + 1, 1) >= 0.02).astype(np.float32)),
        "command": torch.from_numpy(rs.randint(0, 4, size=(T + 1, 1)).astype(np.int32)),
        "hn": torch.zeros(T + 1, feature_dims),
        "cn": torch.zeros(T + 1, feature_dims),
    }
    return st
