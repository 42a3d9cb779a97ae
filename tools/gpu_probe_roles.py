"""Where do the layer1 (tc_flat3x3) and stem (tc_stem_pool) kernels wait? Per-role cycle counters (cadre_debug_clk)
accumulated over one encoder forward of B frames; printed as cycles per tile / conv row, averaged over CTAs."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib  # noqa: E402
from cadre_b200.encoder import Encoder  # noqa: E402
from cadre_b200 import fixtures as R


def persist_table(d, names):
    print("== persistent conv / linear launches in order (cycles per tile, mean over CTAs that worked; CTA pairs: leader CTAs only for the mma rows)")
    print(f"{'launch':28s} {'tiles':>6s} {'prod.wait':>9s} {'mma.tmemE':>9s} {'mma.opsF':>9s} {'mma.loop':>9s} {'epi.bar':>8s} {'epi.tfull':>9s} {'epi.rest':>9s} {'(tmem ld':>9s} {'bar2)':>7s}")
    for r in range(3, d.shape[0]):
        n = d[r, :, 10]
        act = n > 0
        if not act.any():
            continue
        nt = n[act].mean()
        # epilogue counters exist for every CTA; tiles per CTA taken from the mma thread of leader CTAs
        ep = d[r, :, 8] > 0
        per = lambda i, m: (d[r, m, i].sum() / max(n[act].sum(), 1)) * (act.sum() / max(m.sum(), 1))  # noqa: E731
        print(f"{names[r-3] if r-3 < len(names) else r-3!s:28s} {nt:6.1f} {per(1, act):9.0f} {per(2, act):9.0f} {per(3, act):9.0f} {per(5, act):9.0f} "
              f"{per(6, ep):8.0f} {per(7, ep):9.0f} {per(8, ep):9.0f} {per(11, ep):9.0f} {per(12, ep):7.0f}")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    dev = torch.device("cuda:0")
    enc = Encoder(R.danet_fixture_state(0), "cuda:0", max_batch=B)
    xb = torch.rand(B, 4, 144, 256, device=dev)
    out = torch.empty(B, 512, device=dev)
    for _ in range(3):
        enc.forward_f32(xb, out)
    torch.cuda.synchronize()
    dbg = torch.zeros(3 + 24, 148, 16, dtype=torch.int64, device=dev)
    L = _lib.lib()
    L.cadre_debug_clk(ctypes.c_void_p(dbg.data_ptr()))
    enc.forward_f32(xb, out)
    torch.cuda.synchronize()
    L.cadre_debug_clk(ctypes.c_void_p(0))
    d = dbg.cpu().numpy().astype(np.float64)
    names = {0: "stem+pool (per conv row)", 1: "layer1 conv, no residual (per tile)", 2: "layer1 conv + residual (per tile)"}
    for r in range(3):
        n = d[r, :, 10]
        per = lambda i: (d[r, :, i] / np.maximum(n, 1)).mean()  # noqa: E731
        print(f"== {names[r]}: units/CTA {n.mean():.1f}")
        print(f"  producer  wait slot-empty      {per(1):8.0f}")
        print(f"  mma       wait tmem-empty      {per(2):8.0f}")
        print(f"  mma       wait operands-full   {per(3):8.0f}")
        print(f"  mma       issue + commit       {per(4):8.0f}")
        print(f"  mma       loop total           {per(5):8.0f}")
        if r == 0:
            print(f"  epilogue  wait tmem-full       {per(7):8.0f}")
            print(f"  epilogue  conv row -> smem     {per(8):8.0f}")
            print(f"  epilogue  pool + store (avg)   {per(9):8.0f}")
        else:
            print(f"  epilogue  store-read + barrier {per(6):8.0f}")
            print(f"  epilogue  wait tmem-full       {per(7):8.0f}")
            print(f"  epilogue  ld..store issue      {per(8):8.0f}")
    names = ["layer2.0.conv1", "layer2.0.down", "layer2.0.conv2", "layer2.1.conv1", "layer2.1.conv2", "layer3.0.conv1",
             "layer3.0.down", "layer3.0.conv2", "layer3.1.conv1", "layer3.1.conv2", "layer4.0.conv1", "layer4.0.down",
             "layer4.0.conv2", "layer4.1.conv1", "layer4.1.conv2", "conv5a|5c", "pam.value", "conv51", "conv52+sum",
             "fc1"]
    persist_table(d, names)


if __name__ == "__main__":
    main()
