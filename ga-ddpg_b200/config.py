"""Hyper-parameters the update path reads (defaults of /root/reference/experiments/config.py:67-131 and of
experiments/model_spec/rl_pointnet_model_spec.yaml:13-24).  Every RL_TRAIN key becomes an attribute of the
agent in the reference (agent.py:22-23); the fused agent accepts the same mapping and falls back to these."""

DEFAULTS = dict(
    clip_grad=0.5, gamma=0.95, batch_size=256, updates_per_step=4, hidden_size=256, tau=0.0001, lr=3e-4, value_lr=3e-4,
    lr_gamma=0.5, value_lr_gamma=0.5, feature_input_dim=512, ddpg_coefficients=[0.0, 0.0, 1.0, 1.0, 0.2],
    value_milestones=[20000, 40000, 60000, 80000], policy_milestones=[20000, 40000, 60000, 80000],
    mix_milestones=[4000, 8000, 20000, 40000, 60000, 80000, 100000, 140000, 180000],
    mix_policy_ratio_list=[0.1, 0.2], mix_value_ratio_list=[1.0], policy_extra_latent=-1, critic_extra_latent=-1,
    train_value_feature=True, train_feature=True, use_action_limit=True, sa_channel_concat=True, use_time=True,
    value_model=True, shared_feature=False, policy_update_gap=2, policy_aux=True, critic_aux=True, action_noise=0.01,
    noise_ratio_list=[3.0, 2.5, 2.0, 1.5, 1, 0.5], noise_type="uniform", target_update_interval=3000, channel_num=5,
    overwrite_feat_milestone=[], env_name="PandaYCBEnv", concat_option="point_wise", reinit_optim=False, reinit_lr=1e-4,
    # model spec (state_feature_extractor)
    extra_latent=1, feat_lr=1e-3, feat_milestones=[8000, 16000, 30000, 50000, 70000, 90000], feat_gamma=0.3,
)

LOSS_KEYS = ["bc_loss", "policy_grasp_aux_loss", "critic_grasp_aux_loss", "critic_loss", "actor_critic_loss",
             "reward_mask_num", "expert_mask_num", "policy_param", "critic_grad", "critic_param", "train_batch_size"]
