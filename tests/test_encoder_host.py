"""Host-side weight preparation (cadre_b200/encoder.py) checked on CPU against the oracle: BN folding, NHWC
weight layouts, the stem's row-pair layout and the conv8 -> task conv -> Linear fold are exact algebra."""
import numpy as np
import torch
import torch.nn.functional as F

from cadre_b200.encoder import prepare_weights
from oracle import restate as R


def _w(t):
    return t.float()


def test_prepare_weights_matches_oracle_math():
    torch.set_num_threads(8)
    sd = R.danet_fixture_state(0)
    pw = prepare_weights(sd, "cpu")
    tick = R.synthetic_tick(np.random.RandomState(5), seq=2)
    x = torch.from_numpy(R.pre_process(tick["rgb"], tick["route_fig"].copy()))
    with torch.no_grad():
        # --- stem: emulate the kernel's K ordering [row pair j][kw][row r][c] on the padded image
        ref = F.relu(R._bn_eval(F.conv2d(x, sd["backbone.conv1.weight"], sd["backbone.conv1.bias"], 2, 3), sd,
                                "backbone.bn1"))
        xp = torch.zeros(2, 150, 262, 4)
        xp[:, 3:147, 3:259] = x.permute(0, 2, 3, 1)
        P = xp.view(2, 75, 2, 262, 4).permute(0, 1, 3, 2, 4).contiguous()          # [n][y2][x][r][c]
        wk = _w(pw["stem_w"]).view(64, 4, 64)
        for (n, oh, ow) in ((0, 0, 0), (1, 71, 127), (0, 35, 64), (1, 10, 3)):
            acc = _w(pw["stem_b"]).clone()
            for j in range(4):
                a = P[n, oh + j, 2 * ow:2 * ow + 8].reshape(64)
                acc += wk[:, j] @ a
            got = F.relu(acc)
            assert torch.allclose(got, ref[n, :, oh, ow], rtol=2e-2, atol=2e-2)     # bf16 weights
        # --- a folded BasicBlock conv (NHWC-ordered weights back to NCHW) vs conv + BN
        xin = F.max_pool2d(ref, 3, 2, 1)
        w0 = _w(pw["conv_w0"]).view(64, 3, 3, 64).permute(0, 3, 1, 2)
        got = F.conv2d(xin, w0, _w(pw["conv_b0"]), 1, 1)
        want = R._bn_eval(F.conv2d(xin, sd["backbone.layer1.0.conv1.weight"], None, 1, 1), sd,
                          "backbone.layer1.0.bn1")
        assert (got - want).abs().max() < 3e-2 * want.abs().max()
        # --- the linear fold: feat_sum -> fc1 pre-activation for all six heads
        l4 = R.backbone(x, sd)
        sa = R._conv_bn_relu(R.pam(R._conv_bn_relu(l4, sd, "da_head.conv5a"), sd), sd, "da_head.conv51")
        sc = R._conv_bn_relu(R.cam(R._conv_bn_relu(l4, sd, "da_head.conv5c"), sd), sd, "da_head.conv52")
        feat_sum = sa + sc
        da = F.conv2d(feat_sum, sd["da_head.conv8.1.weight"], sd["da_head.conv8.1.bias"])
        a = feat_sum.permute(0, 2, 3, 1).reshape(2, 5120)
        got = a.double() @ pw["fc1_w"].double().t() + pw["fc1_b"].double()
        k = 0
        for task, conv in (("visual", "visual_conv"), ("bc", "bc_conv")):
            t = F.conv2d(da, sd[conv + ".weight"], sd[conv + ".bias"]).reshape(2, -1)
            for role in ("query", "key", "value"):
                p = f"inter_task_att.{task}_{role}_layer.1"
                want = F.linear(t, sd[p + ".weight"], sd[p + ".bias"]).double()
                g = got[:, k * 512:(k + 1) * 512]
                assert (g - want).abs().max() < 2e-2 * want.abs().max() + 1e-3   # bf16-rounded folded weights
                k += 1
    assert pw["fc1_w"].shape == (3072, 5120) and pw["fc2_w"].shape == (6, 256, 512)
    assert pw["head5_w"].shape == (256, 4608) and pw["pam_wqk"].shape == (32, 128)


def test_ingest_chunk_schedule():
    from cadre_b200.ingest import chunk_schedule
    assert chunk_schedule(6400, 640, ramp=(128, 512)) == [128, 512] + [640] * 9     # round 1's e2e ramp
    assert chunk_schedule(6400, 640) == [32, 96, 288] + [640] * 9 + [224]           # default: geometric start-up
    assert chunk_schedule(828, 640) == [32, 96, 288, 412]                # cfg 3, distinct frames only: 4 x 207
    assert chunk_schedule(300, 640) == [300] and chunk_schedule(640, 640) == [640]
    assert chunk_schedule(6400, 640, ramp=()) == [640] * 10
    for n in (1, 127, 641, 5000):
        assert sum(chunk_schedule(n, 640)) == n and max(chunk_schedule(n, 640)) <= 640
