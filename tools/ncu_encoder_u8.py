"""Short uint8-ingest encoder run for ncu: 2 forwards of 640 frames (first = warm-up)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200.encoder import Encoder
from cadre_b200 import fixtures as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
enc = Encoder(R.danet_fixture_state(0), "cuda:0", max_batch=B)
g = torch.Generator().manual_seed(0)
rgb = torch.randint(0, 256, (B, 144, 256, 3), dtype=torch.uint8, generator=g).cuda()
route = (torch.rand(B, 256, 144, generator=g) < 0.1).to(torch.uint8).mul(255).cuda()
meas = torch.rand(B, 3, dtype=torch.float64, generator=g).cuda()
for _ in range(2):
    enc.forward_u8(rgb, route, meas)
torch.cuda.synchronize()
