"""Evaluation with an ensemble of snapshots (reference: eval.py:18-63).

The reference loads `agent_num = len(load_episode)` agents from `<pretrained_path>/models/ppo_model_<episode>.pt`
(eval.py:45-51), lets every agent act on each observation, averages the continuous controls
(`CadreAgent.avg_action`, agent.py:83-95: mean of [steer, throttle, brake]; with more than one agent a mean brake
below 0.5 is released) and steps the environment with the averaged control (eval.py:53-63). The same loop, with the
environment passed in (the CARLA EnvWrapper when a simulator is available, `cadre_b200.synthetic_env.SyntheticEnv`
otherwise: both keep the tick_data contract of env_wrapper.py:857-918)."""
import os


def snapshot_path(pretrained_path, episode):
    """eval.py:49."""
    return os.path.join(pretrained_path, "models", "ppo_model_{}.pt".format(episode))


def load_agent_group(make_agent, pretrained_path, load_episode, device=None):
    """eval.py:45-51: one agent per entry of `load_episode`, each restored from its snapshot.
    `make_agent()` builds a fresh agent (e.g. `lambda: CadreAgent(**agent_cfg)`)."""
    group = []
    for ep in load_episode:
        agent = make_agent()
        agent.load_snapshot(snapshot_path(pretrained_path, ep), device)
        group.append(agent)
    return group


def evaluate(agent_group, env, eval_episode, max_steps=None):
    """eval.py:53-63. Returns one dict per episode: number of steps, summed [steer, throttle] rewards and the
    controls that were applied (the reference only prints where the env wrapper saved its own statistics)."""
    if not agent_group:
        raise ValueError("evaluate() needs at least one agent")
    out = []
    for _ in range(eval_episode):
        obs = env.reset()
        done = False
        steps, ret, controls = 0, [0.0, 0.0], []
        while not done and (max_steps is None or steps < max_steps):
            action_list = []
            for agent in agent_group:
                _, action, *_ = agent.act(obs)
                action_list.append(action)
            control = agent_group[-1].avg_action(action_list)      # eval.py:62 uses the loop's last `agent`
            obs, reward, done, info = env.step(control)
            steps += 1
            ret[0] += float(reward[0])
            ret[1] += float(reward[1])
            controls.append(control)
        out.append({"steps": steps, "return": ret, "controls": controls})
    return out
