"""TEST INFRASTRUCTURE ONLY -- the parity-test cases (BASELINE.json configs, SURVEY.md
section 8d) shared by oracle/make_golden.py and tests/.  Buffer sizes are scaled down
(the full-size rings are exercised by bench.py and the full-size property tests)."""

SAC_KW = dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005, policy_lr=3e-4,
              qf_lr=3e-4, policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3)

CASES = {
    # exp_specs/sac/sac_hopper.yaml:38-47 (script sac_alpha_exp_script.py), B=256 per BASELINE
    "sac_hopper": dict(algo="sac_alpha", obs_dim=11, act_dim=3, batch=256, n_fill=20000,
                       steps=6, sac=dict(SAC_KW, alpha=0.2), seed=11),
    # exp_specs/sac/sac_ant.yaml:38-48 (target_entropy -4)
    "sac_ant": dict(algo="sac_alpha", obs_dim=111, act_dim=8, batch=256, n_fill=8000,
                    steps=3, sac=dict(SAC_KW, alpha=0.2, target_entropy=-4.0), seed=12),
    # yaml batch (512) variant of hopper, fixed alpha branch coverage
    "sac_hopper_b512_fixed_alpha": dict(algo="sac_alpha", obs_dim=11, act_dim=3, batch=512,
                                        n_fill=20000, steps=3,
                                        sac=dict(SAC_KW, alpha=0.2, train_alpha=False), seed=13),
    # run_scripts/sac_exp_script.py -> sac.py (V variant)
    "sacv_hopper": dict(algo="sac_v", obs_dim=11, act_dim=3, batch=256, n_fill=20000, steps=4,
                        sac=dict(SAC_KW, vf_lr=3e-4, alpha=1.0), seed=14),
    # exp_specs/td3/td3_humanoid.yaml:13-14,44-50; B=1024 per BASELINE config 4
    "td3_humanoid": dict(algo="td3", obs_dim=376, act_dim=17, batch=1024, n_fill=6000, steps=4,
                         td3=dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005,
                                  policy_lr=3e-4, qf_lr=3e-4, policy_and_target_update_period=2),
                         policy_noise=0.2, policy_noise_clip=0.5, seed=15),
    "td3_hopper": dict(algo="td3", obs_dim=11, act_dim=3, batch=256, n_fill=20000, steps=5,
                       td3=dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005,
                                policy_lr=3e-4, qf_lr=3e-4, policy_and_target_update_period=2),
                       policy_noise=0.2, policy_noise_clip=0.5, seed=16),
    # exp_specs/her/her_reach_td3.yaml -> her/td3.py: obs_dim = observation (10) + desired_goal (3); policy_lr 6e-4
    "her_td3_reach": dict(algo="td3", obs_dim=13, act_dim=4, batch=256, n_fill=20000, steps=5,
                          her=dict(goal_dim=3, sigma=0.2),
                          td3=dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005, policy_lr=6e-4,
                                   qf_lr=3e-4, policy_and_target_update_period=2),
                          policy_noise=0.2, policy_noise_clip=0.5, seed=25),
    # run_scripts/her_sac_exp_script.py -> her/sac.py: sac_alpha on cat(obs, goal), target entropy -A (her/sac.py:52)
    "her_sac_reach": dict(algo="sac_alpha", obs_dim=13, act_dim=4, batch=256, n_fill=20000, steps=4,
                          her=dict(goal_dim=3), sac=dict(SAC_KW, alpha=0.2, target_entropy=-4.0), seed=27),
    # the yamls' net_size 300 (not a multiple of the 32-wide tile) and a narrow return clip
    "her_td3_h300": dict(algo="td3", obs_dim=13, act_dim=4, batch=128, n_fill=4000, steps=4, hidden=300,
                         her=dict(goal_dim=3, sigma=0.3, clip_return_l=-0.05, clip_return_r=0.02),
                         td3=dict(reward_scale=1.0, discount=0.98, soft_target_tau=0.005, policy_lr=6e-4,
                                  qf_lr=3e-4, policy_and_target_update_period=2),
                         policy_noise=0.2, policy_noise_clip=0.5, seed=26),
    # exp_specs/gail/gail_walker.yaml (gail2, grad_pen 8, reward_scale 2, beta_1 0.25)
    "gail_walker": dict(algo="adv_irl", obs_dim=17, act_dim=6, batch=256, n_fill=20000,
                        n_expert=4000, steps=4, mode="gail2",
                        disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True,
                                  grad_pen_weight=8.0),
                        sac=dict(SAC_KW, reward_scale=2.0, beta_1=0.25, alpha=0.2), seed=17),
    "airl_hopper_clip": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000,
                             n_expert=2000, steps=3, mode="airl", rew_clip_min=-2.0,
                             rew_clip_max=2.0,
                             disc=dict(disc_lr=3e-4, disc_momentum=0.0, use_grad_pen=True,
                                       grad_pen_weight=10.0),
                             sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=18),
    "fairl_hopper_nogp": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000,
                              n_expert=2000, steps=3, mode="fairl",
                              disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=False,
                                        grad_pen_weight=10.0),
                              sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=19),
    "gail_hopper": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000,
                        n_expert=2000, steps=3, mode="gail",
                        disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True,
                                  grad_pen_weight=4.0),
                        sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=20),
    # adv_irl.py:139-179,265-269 state_only (AIRL on cat(s, s')) and :239-255 expert transitions mixed into the policy batch
    "airl_state_only": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000, n_expert=2000, steps=3,
                            mode="airl", state_only=True,
                            disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=10.0),
                            sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=28),
    "gail_expert_mix": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000, n_expert=2000, steps=3,
                            mode="gail", from_expert=32,
                            disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=4.0),
                            sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=29),
    # MLPDisc(hid_act="relu", use_bn=False) (simple_disc_models.py:19-24; the class default activation): no act'' terms in the
    # penalty's double backward -- with and without the penalty, and on ragged shapes
    "gail_hopper_relu": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000, n_expert=2000, steps=3,
                             mode="gail", disc_act="relu",
                             disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=4.0),
                             sac=dict(SAC_KW, reward_scale=1.0, alpha=0.2), seed=30),
    "gail2_relu_nogp_ragged": dict(algo="adv_irl", obs_dim=5, act_dim=2, batch=40, n_fill=400, n_expert=90, steps=3,
                                   mode="gail2", hidden=64, disc_hid=48, disc_act="relu",
                                   disc=dict(disc_lr=3e-4, disc_momentum=0.0, use_grad_pen=False, grad_pen_weight=10.0),
                                   sac=dict(SAC_KW, reward_scale=2.0, alpha=0.2), seed=31),
    # ragged shapes: batch not a multiple of the 32-row tile / of 4, odd hidden widths, tiny dims
    "sac_ragged": dict(algo="sac_alpha", obs_dim=5, act_dim=2, batch=70, n_fill=300, steps=3, hidden=64,
                       sac=dict(SAC_KW, alpha=0.2), seed=21),
    "td3_ragged": dict(algo="td3", obs_dim=7, act_dim=3, batch=72, n_fill=500, steps=4, hidden=96,
                       td3=dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005, policy_lr=3e-4,
                                qf_lr=3e-4, policy_and_target_update_period=2),
                       policy_noise=0.2, policy_noise_clip=0.5, seed=22),
    "gail_ragged": dict(algo="adv_irl", obs_dim=5, act_dim=2, batch=40, n_fill=400, n_expert=90, steps=3,
                        mode="gail2", hidden=64, disc_hid=48,
                        disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=8.0),
                        sac=dict(SAC_KW, reward_scale=2.0, beta_1=0.25, alpha=0.2), seed=23),
}

# adv_irl.py:126-131 with num_disc_updates_per_loop_iter / num_policy_updates_per_loop_iter != 1 (gail_humanoid.yaml uses
# 100/100): one case "step" = one _do_training call = n_disc discriminator updates, then n_policy policy updates
LOOP_CASES = {
    "gail_nd2_np3": dict(algo="adv_irl", obs_dim=11, act_dim=3, batch=128, n_fill=5000, n_expert=2000, steps=2,
                         n_disc=2, n_policy=3, mode="gail",
                         disc=dict(disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=4.0),
                         sac=dict(SAC_KW, reward_scale=2.0, alpha=0.2), seed=24),
}

# HER relabel-at-sample (relabel_replay_buffer.py:63-131) feeding her/td3.py: observation 10 + goal 3, 12 episodes of 50 steps
HER_RELABEL_CASES = {
    "her_td3_relabel": dict(algo="td3", obs_dim=13, act_dim=4, batch=64, n_fill=800, steps=4,
                            her=dict(goal_dim=3, sigma=0.2), her_ratio=0.8, threshold=0.05, n_episodes=12, T=50,
                            td3=dict(reward_scale=1.0, discount=0.98, soft_target_tau=0.005, policy_lr=6e-4,
                                     qf_lr=3e-4, policy_and_target_update_period=2),
                            policy_noise=0.2, policy_noise_clip=0.5, seed=30),
    # the same buffer feeding her/sac.py (the SAC program gathers the NEXT step's batch in its last phase)
    "her_sac_relabel": dict(algo="sac_alpha", obs_dim=13, act_dim=4, batch=64, n_fill=800, steps=4,
                            her=dict(goal_dim=3), her_ratio=0.8, threshold=0.05, n_episodes=12, T=50,
                            sac=dict(SAC_KW, alpha=0.2, target_entropy=-4.0), seed=31),
}

HIDDEN = (256, 256)
DISC_HID = 128
BUFFER_SEED = 1      # policy replay buffer index RNG
EXPERT_SEED = 3      # expert replay buffer index RNG
DATA_SEED = 7
EXPERT_DATA_SEED = 8
EPS_SEED0 = 1000     # torch.manual_seed(EPS_SEED0 + t) before step t
