"""ctypes mirror of include/ilswiss_b200.h (the C ABI of libilswiss_b200.so).

Pure declarations: importing this module does not load the library (see _lib.py)."""
import ctypes as C

ABI_VERSION = 1
UPDATE_BOTH, UPDATE_DISC_ONLY, UPDATE_POLICY_ONLY = 0, 1, 2
LOSS_SLOTS = 16
IPC_HANDLE_BYTES = 64

ALGO_SAC_ALPHA, ALGO_TD3, ALGO_SAC_V = 1, 2, 3
DISC_MODES = {"airl": 0, "gail": 1, "gail2": 2, "fairl": 3}
DISC_ACTS = {"tanh": 0, "relu": 1}     # ilsw_disc_act

(L_QF1, L_QF2, L_POLICY, L_ALPHA_LOSS, L_ALPHA, L_VF, L_DISC_CE, L_DISC_ACC, L_GRAD_PEN,
 L_REW_MEAN, L_REW_STD, L_REW_MAX, L_REW_MIN, L_Q1_MEAN, L_LOGPI_MEAN, L_QT_MEAN) = range(16)

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class Mlp(C.Structure):
    _fields_ = [("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("in_dim", C.c_int), ("hidden", C.c_int), ("out_dim", C.c_int),
                ("log_std_head", C.c_int)]


class TrainerConfig(C.Structure):
    _fields_ = [("algo", C.c_int), ("obs_dim", C.c_int), ("act_dim", C.c_int), ("batch", C.c_int),
                ("max_steps_per_call", C.c_int), ("gemm_precision", C.c_int),
                ("reward_scale", C.c_double), ("discount", C.c_double), ("soft_target_tau", C.c_double),
                ("policy_lr", C.c_double), ("qf_lr", C.c_double), ("vf_lr", C.c_double), ("alpha_lr", C.c_double),
                ("beta_1", C.c_double), ("beta_2", C.c_double), ("adam_eps", C.c_double),
                ("alpha", C.c_double), ("train_alpha", C.c_int), ("target_entropy", C.c_double),
                ("policy_mean_reg_weight", C.c_double), ("policy_std_reg_weight", C.c_double),
                ("policy_and_target_update_period", C.c_int),
                ("policy_noise", C.c_double), ("policy_noise_clip", C.c_double), ("max_act", C.c_double),
                ("her", C.c_int), ("her_sigma", C.c_double), ("min_act", C.c_double),
                ("clip_return_l", C.c_double), ("clip_return_r", C.c_double)]


class DiscConfig(C.Structure):
    _fields_ = [("mode", C.c_int), ("batch", C.c_int), ("disc_lr", C.c_double), ("disc_momentum", C.c_double),
                ("use_grad_pen", C.c_int), ("grad_pen_weight", C.c_double), ("clamp_magnitude", C.c_double),
                ("rew_clip_min_on", C.c_int), ("rew_clip_max_on", C.c_int),
                ("rew_clip_min", C.c_double), ("rew_clip_max", C.c_double),
                ("state_only", C.c_int), ("policy_batch_from_expert", C.c_int), ("hid_act", C.c_int)]


class HerSamplingDesc(C.Structure):
    _fields_ = [("enabled", C.c_int), ("n_traj", C.c_int), ("traj_start", C.c_void_p), ("traj_len", C.c_void_p),
                ("next_achieved_goal", C.c_void_p), ("goal_dim", C.c_int), ("relabel_num", C.c_int),
                ("distance_threshold", C.c_float), ("inj_idx_her", C.c_void_p)]


class Inject(C.Structure):
    _fields_ = [("idx", C.c_void_p), ("eps_next", C.c_void_p), ("eps_cur", C.c_void_p),
                ("idx_expert", C.c_void_p), ("idx_policy_d", C.c_void_p), ("gp_eps", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("rew", C.c_void_p), ("term", C.c_void_p),
                ("next_obs", C.c_void_p)]


class State(C.Structure):
    _fields_ = [("log_alpha", C.c_double), ("alpha_exp_avg", C.c_double), ("alpha_exp_avg_sq", C.c_double),
                ("alpha_step", C.c_int), ("adam_step", C.c_int * 8), ("n_train_steps_total", C.c_int)]


# name -> (restype, argtypes); every symbol include/ilswiss_b200.h declares
PROTOTYPES = {
    "ilsw_abi_version": (C.c_int, []),
    "ilsw_last_error": (C.c_char_p, []),
    "ilsw_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "ilsw_rb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_int]),
    "ilsw_rb_destroy": (C.c_int, [C.c_void_p]),
    "ilsw_rb_host_row_floats": (C.c_int, [C.c_void_p]),
    "ilsw_rb_row_stride": (C.c_int, [C.c_void_p]),
    "ilsw_rb_capacity": (C.c_int64, [C.c_void_p]),
    "ilsw_rb_size": (C.c_int64, [C.c_void_p]),
    "ilsw_rb_top": (C.c_int64, [C.c_void_p]),
    "ilsw_rb_rows_ptr": (C.c_void_p, [C.c_void_p]),
    "ilsw_rb_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ilsw_rb_commit": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilsw_rb_load_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ilsw_rb_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ilsw_rb_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ilsw_rb_clear": (C.c_int, [C.c_void_p]),
    "ilsw_rb_set_cursor": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "ilsw_mlp_num_params": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ilsw_trainer_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(TrainerConfig), C.POINTER(Mlp), C.c_int]),
    "ilsw_trainer_attach_disc": (C.c_int, [C.c_void_p, C.POINTER(DiscConfig), C.POINTER(Mlp)]),
    "ilsw_trainer_set_update_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "ilsw_trainer_set_her": (C.c_int, [C.c_void_p, C.POINTER(HerSamplingDesc)]),
    "ilsw_trainer_destroy": (C.c_int, [C.c_void_p]),
    "ilsw_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Inject), C.POINTER(Batch),
                             C.c_uint64, C.c_int, C.c_void_p]),
    "ilsw_read_losses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ilsw_read_losses_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ilsw_check_abort": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilsw_stats_floats": (C.c_int, [C.c_void_p]),
    "ilsw_read_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ilsw_get_state": (C.c_int, [C.c_void_p, C.POINTER(State), C.c_void_p]),
    "ilsw_set_state": (C.c_int, [C.c_void_p, C.POINTER(State), C.c_void_p]),
    "ilsw_describe_program": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ilsw_num_phases": (C.c_int, [C.c_void_p]),
    "ilsw_read_phase_ns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ilsw_read_tile_ns": (C.c_int, [C.c_void_p]),
    "ilsw_trainer_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ilsw_read_cta_ns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ilsw_kernel_launches": (C.c_int64, [C.c_void_p]),
    "ilsw_trainer_uses_tc5": (C.c_int, [C.c_void_p]),
    "ilsw_policy_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ilsw_policy_act_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ilsw_replica_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ilsw_replica_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ilsw_replica_buffer_bytes": (C.c_int64, [C.c_void_p]),
    "ilsw_replica_connect_symm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.c_int64]),
}


def declare(lib):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
