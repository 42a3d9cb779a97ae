"""Short encoder run for ncu: 2 forwards of 640 frames (first = warm-up)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200.encoder import Encoder
from cadre_b200 import fixtures as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
enc = Encoder(R.danet_fixture_state(0), "cuda:0", max_batch=B)
x = torch.rand(B, 4, 144, 256, device="cuda")
for _ in range(2):
    enc.forward_f32(x)
torch.cuda.synchronize()
