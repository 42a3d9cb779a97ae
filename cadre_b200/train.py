"""Worker loop (reference: ppo_agent/train.py:14-127) against the synthetic environment, with the
parameter-server handshake (train.py:101-110 + chief.py) replaced by `Learner.update_step` (batched local
workers + NCCL all-reduce + replicated clip/Adam). `main()` is the per-GPU launcher that replaces main.py:25-72:
one process per GPU (torchrun), each stepping `workers_per_gpu` logical workers."""
import os

import numpy as np
import torch

from .actor import BatchedActor
from .agent import CadreAgent
from .config import load_config
from .learner import Learner, RolloutPool
from .models import get_vae_output, init_state
from .synthetic_env import SyntheticEnv


def train(rank, train_cfg, agent_cfg, env_cfg, rollout_cfg, danet_state, ppo_state=None, workers=1, max_episode=None,
          process_group=None, log=print, batched_acting=True):
    """One rank: `workers` logical workers (env + rollout pair each) sharing one agent replica. With
    `batched_acting` the rank's workers are stepped together through `BatchedActor` (one encoder batch of the newest
    frames + one routed LSTM / head evaluation per tick) instead of one `agent.act` call per worker."""
    device = torch.device("cuda:" + str(agent_cfg.model_cfg.device_num))
    envs = [SyntheticEnv(dict(env_cfg, rank=rank * workers + w, seq_length=rollout_cfg.seq_length))
            for w in range(workers)]
    hidden_size, _ = get_vae_output(agent_cfg.model_cfg)
    ppo_state = ppo_state if ppo_state is not None else init_state(seed=0)
    num_steps = rollout_cfg.num_steps
    mb = num_steps // rollout_cfg.mini_batch_num
    agent = CadreAgent(**dict(agent_cfg, rank=rank), danet_state=danet_state, ppo_state=ppo_state, mini_batch=mb)
    learner = Learner(workers, mb, ppo_state, device, agent_cfg.clip, agent_cfg.value_coeff, agent_cfg.clip_coeff,
                      agent_cfg.ent_coeff, train_cfg.lr, train_cfg.max_grad_norm, process_group,
                      seeds=[env_cfg.get("seed", 0) + rank * workers + w for w in range(workers)])
    agent.owner.params = learner.params          # the acting replica reads the learner's parameters directly
    pool = RolloutPool(workers, dict(rollout_cfg, hidden_size=hidden_size), device)
    obs = [e.reset() for e in envs]
    done = [False] * workers
    history = []
    actor = BatchedActor(agent, workers) if batched_acting else None
    for episode in range(max_episode if max_episode is not None else train_cfg.max_episode):
        for _ in range(num_steps):                                        # train.py:55-74
            acted = actor.act(obs) if actor is not None else None
            for w, env in enumerate(envs):
                command = obs[w]["command"]
                feat, action, logp, value, hidden = acted[w] if acted is not None else agent.act(obs[w])
                obs[w], reward, done[w], info = env.step(agent.convert_action(action))
                for h in range(2):
                    mask = torch.tensor([[0.0] if info["action_done"][h] else [1.0]])
                    pool.storages[w][h].insert(feat, action[h], logp[h], value[h], reward[h], mask, hidden, command)
                if done[w]:
                    obs[w] = env.reset()
                    if actor is not None:
                        actor.reset(w)
        nv = torch.zeros(workers, 2)
        for w in range(workers):                                          # train.py:76-79
            vs, vt = agent.get_value(done[w], pool.storages[w][0].get_last(), pool.storages[w][1].get_last())
            nv[w, 0], nv[w, 1] = float(vs.reshape(-1)[0]), float(vt.reshape(-1)[0])
        pool.compute_returns(nv, normalize=train_cfg.use_adv_norm)        # train.py:81-88
        learner.learn(pool, train_cfg.ppo_epoch)                          # train.py:93-110
        L = learner.scaled_losses(mean=True).mean(0)     # mean over update steps and workers
        history.append(L.tolist())
        if episode % train_cfg.log_interval == 0 and rank == 0:
            log("Episode: {}, value loss: {:.4f}, policy loss: {:.4f}, entropy loss: {:.4f}".format(episode, *L))
    return learner, history


def main():
    cfg = load_config(os.environ.get("CADRE_CONFIG") or None) if os.environ.get("CADRE_CONFIG") else load_config()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_ALGO", "Ring")   # faster than NCCL's default for the 36 - 78 MB gradient ranges (8 x B200)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg.agent_cfg.model_cfg.device_num = cfg.agent_cfg.model_cfg.vae_device = local
    ckpt = os.environ.get("CADRE_ENCODER_CKPT")
    if not ckpt:
        raise SystemExit("set CADRE_ENCODER_CKPT to the reference perception checkpoint ({'autoencoder': state_dict})")
    danet_state = torch.load(ckpt, map_location="cpu")["autoencoder"]
    # env_cfg.num_processes logical workers in total (main.py:41-60), split evenly over the ranks. Gradients are SUMMED
    # over workers like the reference (models.py:231-239), so the effective step (clip at max_grad_norm, Adam's eps)
    # depends on the total worker count: it must match the configured one, not grow with the number of GPUs.
    if cfg.env_cfg.num_processes % world:
        raise SystemExit(f"env_cfg.num_processes={cfg.env_cfg.num_processes} is not divisible by WORLD_SIZE={world}")
    workers = cfg.env_cfg.num_processes // world
    train(rank, cfg.train_cfg, cfg.agent_cfg, cfg.env_cfg, cfg.rollout_cfg, danet_state, workers=workers)


if __name__ == "__main__":
    main()
