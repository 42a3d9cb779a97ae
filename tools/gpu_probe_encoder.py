"""GPU probe: per-stage encoder error vs the oracle, and an encoder batch sweep timing."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200.encoder import Encoder  # noqa: E402
from oracle import restate as R  # noqa: E402


def rel(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def main():
    torch.set_num_threads(8)
    sd = R.danet_fixture_state(0)
    enc = Encoder(sd, "cuda:0", max_batch=1024)
    B = 5
    x = torch.from_numpy(np.random.RandomState(0).rand(B, 4, 144, 256).astype(np.float32))
    lat = enc.forward_f32(x.cuda()).cpu()
    with torch.no_grad():
        st = F.relu(R._bn_eval(F.conv2d(x, sd["backbone.conv1.weight"], sd["backbone.conv1.bias"], 2, 3), sd,
                               "backbone.bn1"))
        l4 = R.backbone(x, sd)
        f1 = R._conv_bn_relu(l4, sd, "da_head.conv5a")
        f2 = R._conv_bn_relu(l4, sd, "da_head.conv5c")
        sa, sc = R.pam(f1, sd), R.cam(f2, sd)
        fs = R._conv_bn_relu(sa, sd, "da_head.conv51") + R._conv_bn_relu(sc, sd, "da_head.conv52")
        ref = R.encoder_latent(x, sd)

    def nhwc(buf, h, w, c):
        return enc.debug_buffer(buf, B).view(B, h, w, c).float().cpu().permute(0, 3, 1, 2)
    print("stem   rel", rel(nhwc(5, 72, 128, 64), st))
    print("l4     rel", rel(nhwc(0, 5, 8, 512), l4))
    h5 = nhwc(2, 5, 8, 256)
    print("conv5a rel", rel(h5[:, :128], f1), "conv5c rel", rel(h5[:, 128:], f2))
    print("pam    rel", rel(nhwc(3, 5, 8, 128), sa), "cam rel", rel(nhwc(4, 5, 8, 128), sc))
    print("fsum   rel", rel(nhwc(1, 5, 8, 128), fs))
    print("latent rel", rel(lat, ref), "launches", enc.launches_per_forward, flush=True)

    for Bs in (1, 8, 64, 256, 1024):
        xb = torch.rand(Bs, 4, 144, 256, device="cuda")
        out = torch.empty(Bs, 512, device="cuda")
        for _ in range(3):
            enc.forward_f32(xb, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        it = 10
        e0.record()
        for _ in range(it):
            enc.forward_f32(xb, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"encoder B={Bs}: {ms:.3f} ms  {Bs/ms*1e3:.0f} frames/s  {3.0875*Bs/ms:.1f} TFLOP/s-equiv", flush=True)


if __name__ == "__main__":
    main()
