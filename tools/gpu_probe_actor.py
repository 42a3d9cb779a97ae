"""Acting throughput: E environments per tick through BatchedActor (frame cache, one batch) against E sequential
CadreAgent.act calls on the full 8-frame windows."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import fixtures as R
from cadre_b200.agent import CadreAgent
from cadre_b200.actor import BatchedActor
from cadre_b200.config import load_config
cfg = load_config()
agent = CadreAgent(**cfg.agent_cfg, danet_state=R.danet_fixture_state(0), ppo_state=R.ppo_fixture_state(0), max_encoder_batch=512)
for E in (8, 64):
    rs = np.random.RandomState(0)
    n = 8 + 12
    streams = [dict(rgb=rs.randint(0, 256, size=(n, 144, 256, 3)).astype(np.uint8),
                    route_fig=(rs.rand(n, 256, 144) < 0.1).astype(np.uint8) * 255, measurements=rs.rand(n, 3)) for _ in range(E)]
    def ticks_at(t):
        return [dict(rgb=s["rgb"][t:t + 8], route_fig=s["route_fig"][t:t + 8], measurements=s["measurements"][t:t + 8],
                     command=int((t + e) % 4)) for e, s in enumerate(streams)]
    actor = BatchedActor(agent, E)
    actor.act(ticks_at(0)); actor.act(ticks_at(1)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(2, 12):
        actor.act_batch(ticks_at(t))
    torch.cuda.synchronize(); tb = (time.perf_counter() - t0) / 10
    actor_nc = BatchedActor(agent, E, verify_window=False)
    actor_nc.act(ticks_at(0)); actor_nc.act(ticks_at(1)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(2, 12):
        actor_nc.act_batch(ticks_at(t))
    torch.cuda.synchronize(); tn = (time.perf_counter() - t0) / 10
    tk = ticks_at(5)
    for e in range(min(E, 4)): agent.act(dict(tk[e]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for e in range(E): agent.act(dict(tk[e]))
    torch.cuda.synchronize(); ts = time.perf_counter() - t0
    print(f"E={E}: BatchedActor {tb*1e3:.2f} ms/tick ({E/tb:.0f} env-steps/s; {tn*1e3:.2f} ms without the host window check), "
          f"{E} x CadreAgent.act {ts*1e3:.2f} ms/tick ({E/ts:.0f} env-steps/s) -> {ts/tb:.1f}x", flush=True)
