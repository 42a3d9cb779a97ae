"""Default hyper-parameters of the learner path. Same names and values as the reference's
config_files/agent_config.py:1-58,74,96 (the CARLA-only env_cfg entries are dropped because the synthetic
environment does not use them). Loaded by cadre_b200.config.load_config."""
THROTTLE_CONTROL = {0: [0, 0], 1: [0, 1], 2: [0.6, 0]}

STEER_CONTROL = {i: (i - 8) / 16. for i in range(17)}
STEER_CONTROL.update({17: 9. / 16, 18: -9. / 16, 19: 10. / 16, 20: -10. / 16, 21: 11. / 16, 22: -11. / 16,
                      23: 12. / 16., 24: -12. / 16., 25: 13. / 16, 26: -13. / 16., 27: 14. / 16., 28: -14. / 16.,
                      29: 15. / 16., 30: -15. / 16., 31: 1., 32: -1.})

rollout_cfg = dict(num_steps=200, mini_batch_num=2, feature_dims=512 + 18, seq_length=8, use_gae=True, gamma=0.99,
                   tau=0.95)

agent_cfg = dict(
    rank=-1,
    model_cfg=dict(use_lstm=True, vae_device=0, device_num=0, vae_params="CoPM", measurement_dim=18,
                   num_output=dict(steer=len(STEER_CONTROL), throttle=len(THROTTLE_CONTROL)), command_num=4),
    frame=8, STEER_CONTROL=STEER_CONTROL, THROTTLE_CONTROL=THROTTLE_CONTROL,
    ent_coeff=0.01, value_coeff=0.1, clip_coeff=1., clip=0.1)

train_cfg = dict(max_episode=3000, max_grad_norm=250, use_adv_norm=True, ppo_epoch=4, lr=3e-4, save_interval=100,
                 log_interval=10)

env_cfg = dict(root_path="result", num_processes=4, seq_length=8, width=256, height=144, seed=0,
               done_prob=0.005, action_done_prob=0.02)
