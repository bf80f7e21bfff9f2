"""ctypes mirrors of the engine structs / entry points of include/gemmgan.h."""
from __future__ import annotations

import ctypes as C

VARIANT_VANILLA, VARIANT_FILM, VARIANT_PAPER, VARIANT_CROSS, VARIANT_CONCAT, VARIANT_IMG, VARIANT_LABEL = 0, 1, 2, 3, 4, 5, 6
VARIANT_ATTN = 7   # conditional_gan_attention.py
OPT_RMSPROP, OPT_ADAM, OPT_ADAMW = 0, 1, 2
NET_GEN, NET_DISC = 0, 1
PHASE_STAGE0, PHASE_NO_JOIN = 16, 64
TRAIN_GEN_EVAL = 2   # GG_TRAIN_GEN_EVAL: flag of gg_engine_disc_grads' `training` argument

# enum gg_param_slot
P_FILM_W, P_FILM_B, P_TEXT_W, P_TEXT_B, P_PATCH_W, P_PATCH_B, P_CLS = range(7)
P_LAYER0 = 7
P_PENC_LN_W, P_PENC_LN_B = P_FILM_W, P_FILM_B   # GG_VARIANT_IMG: patch-encoder LayerNorm vectors
P_EMB0, P_EMB1 = P_TEXT_W, P_PATCH_W             # GG_VARIANT_LABEL: the two embedding tables
P_BN_W, P_BN_B = P_FILM_W, P_FILM_B              # GG_VARIANT_ATTN (generator): BatchNorm1d weight / bias
(L_IN_W, L_IN_B, L_OUT_W, L_OUT_B, L_FF1_W, L_FF1_B, L_FF2_W, L_FF2_B,
 L_N1_W, L_N1_B, L_N2_W, L_N2_B) = range(12)
L_COUNT = 12
P_P2T_IN_W = P_LAYER0 + 24
P_P2T_IN_B, P_P2T_OUT_W, P_P2T_OUT_B = P_P2T_IN_W + 1, P_P2T_IN_W + 2, P_P2T_IN_W + 3
P_T2P_IN_W, P_T2P_IN_B, P_T2P_OUT_W, P_T2P_OUT_B = (P_P2T_IN_W + 4, P_P2T_IN_W + 5, P_P2T_IN_W + 6,
                                                    P_P2T_IN_W + 7)
P_TR0_W, P_TR0_B, P_TR1_W, P_TR1_B, P_FIN_W, P_FIN_B = (P_P2T_IN_W + 8 + i for i in range(6))
NSLOTS = P_FIN_B + 1

STAT_LOSS_REAL, STAT_LOSS_FAKE, STAT_GP, STAT_D_LOSS, STAT_G_LOSS = 0, 1, 2, 3, 4
STAT_D_GRAD_NORM, STAT_D_CLIP_COEF, STAT_G_GRAD_NORM, STAT_G_CLIP_COEF = 5, 6, 7, 8
STATS_COUNT = 16


class ModelCfg(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("B", C.c_int32), ("G", C.c_int32), ("L", C.c_int32), ("E", C.c_int32),
        ("H", C.c_int32), ("Dt", C.c_int32), ("Dp", C.c_int32), ("P", C.c_int32), ("T", C.c_int32),
        ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("ffn", C.c_int32), ("tower_bias", C.c_int32),
        ("slope", C.c_float), ("dropout_p", C.c_float), ("gp_weight", C.c_float), ("clip_d", C.c_float),
        ("clip_g", C.c_float), ("ln_eps", C.c_float), ("optimizer", C.c_int32), ("gemm_impl", C.c_int32),
        ("seed", C.c_uint64),
    ]


class NetBuffers(C.Structure):
    _fields_ = [
        ("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("step_count", C.c_void_p),
        ("n_used", C.c_int64), ("off", C.c_int64 * NSLOTS),
    ]


class WgradItem(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("ld_dy", C.c_int64), ("x", C.c_void_p), ("ld_x", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("out", C.c_void_p), ("ld", C.c_int64),
                ("bias", C.c_void_p)]


class AttnArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int64), ("q_mod", C.c_int32),
                ("k", C.c_void_p), ("v", C.c_void_p), ("ldkv", C.c_int64), ("kv_mod", C.c_int32),
                ("mask", C.c_void_p), ("mask_mod", C.c_int32),
                ("nb", C.c_int32), ("H", C.c_int32), ("hd", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
                ("drop_p", C.c_float), ("rng", C.c_void_p), ("site", C.c_uint32),
                ("o", C.c_void_p), ("ldo", C.c_int64),
                ("dout", C.c_void_p), ("lddo", C.c_int64), ("dq", C.c_void_p), ("lddq", C.c_int64),
                ("dk", C.c_void_p), ("dv", C.c_void_p), ("lddkv", C.c_int64), ("stat", C.c_void_p),
                ("dbits", C.c_void_p)]


class EncLayerParams(C.Structure):
    """gg_enc_layer_params (include/gemmgan.h)."""
    _fields_ = [("nb", C.c_int32), ("S", C.c_int32), ("E", C.c_int32), ("F", C.c_int32), ("n_heads", C.c_int32),
                ("save_rows", C.c_int64), ("x", C.c_void_p),
                ("w_in", C.c_void_p), ("ld_in", C.c_int64), ("w_out", C.c_void_p), ("ld_out", C.c_int64),
                ("w_ff1", C.c_void_p), ("ld_ff1", C.c_int64), ("w_ff2", C.c_void_p), ("ld_ff2", C.c_int64),
                ("b_in", C.c_void_p), ("b_out", C.c_void_p), ("b_ff1", C.c_void_p), ("b_ff2", C.c_void_p),
                ("g1", C.c_void_p), ("be1", C.c_void_p), ("g2", C.c_void_p), ("be2", C.c_void_p),
                ("mask", C.c_void_p), ("mask_mod", C.c_int32), ("drop_p", C.c_float), ("eps", C.c_float),
                ("rng", C.c_void_p), ("site", C.c_uint32),
                ("qkv", C.c_void_p), ("ao", C.c_void_p), ("z1", C.c_void_p), ("x1", C.c_void_p), ("h", C.c_void_p),
                ("z2", C.c_void_p), ("out", C.c_void_p),
                ("mean1", C.c_void_p), ("rstd1", C.c_void_p), ("mean2", C.c_void_p), ("rstd2", C.c_void_p),
                ("dbits1", C.c_void_p), ("dbits2", C.c_void_p), ("dbits3", C.c_void_p)]


class EncFfnBwdParams(C.Structure):
    """gg_enc_ffn_bwd_params (include/gemmgan.h)."""
    _fields_ = [("rows", C.c_int64), ("dout", C.c_void_p), ("z2", C.c_void_p), ("mean2", C.c_void_p),
                ("rstd2", C.c_void_p), ("gamma2", C.c_void_p), ("h", C.c_void_p), ("w2t", C.c_void_p),
                ("ld_w2t", C.c_int64), ("w1t", C.c_void_p), ("ld_w1t", C.c_int64), ("drop_p", C.c_float),
                ("rng", C.c_void_p), ("site", C.c_uint32), ("gh", C.c_void_p), ("gb", C.c_void_p), ("dbits", C.c_void_p)]


class ColsumItem(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("ld", C.c_int64), ("rows", C.c_int64), ("N", C.c_int32), ("out", C.c_void_p)]


def declare_evalmetrics(L: C.CDLL) -> None:
    """Entry points of csrc/evalmetrics.cu (kept apart so that tests can bind the host-emulated build of that file)."""
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.gg_pairwise_distance.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp, i64, vp]
    L.gg_row_kth_smallest.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp]
    L.gg_row_membership.argtypes = [vp, i64, i32, i32, vp, i32, f32, vp, vp, vp, vp, vp]
    L.gg_col_hits.argtypes = [vp, i64, i32, i32, vp, i32, vp, vp]
    L.gg_standardize_columns.argtypes = [vp, i64, i32, i32, vp, i64, vp]
    L.gg_gene_correlation.argtypes = [vp, i64, vp, i64, i32, i32, i32, vp, i64, vp]
    L.gg_gamma_moments_workspace_bytes.argtypes = [i32]
    L.gg_gamma_moments_workspace_bytes.restype = i64
    L.gg_gamma_moments.argtypes = [vp, i64, i32, vp, i64, i32, i32, vp, i64, vp, vp]


def declare(L: C.CDLL) -> None:
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.gg_engine_workspace_bytes.argtypes = [C.POINTER(ModelCfg), C.POINTER(i64)]
    L.gg_engine_create.argtypes = [C.POINTER(ModelCfg), C.POINTER(NetBuffers), C.POINTER(NetBuffers), vp, i64, vp,
                                   C.POINTER(vp)]
    L.gg_engine_destroy.argtypes = [vp]
    L.gg_engine_destroy.restype = None
    L.gg_engine_refresh_shadows.argtypes = [vp, i32, vp]
    L.gg_engine_set_lanes.argtypes = [vp, i32]
    L.gg_engine_set_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.gg_engine_set_labels.argtypes = [vp, vp, vp, vp]
    L.gg_film_patch_encode.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.gg_gemm_layernorm.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, f32, f32, vp, C.c_uint32, vp]
    L.gg_xw_f32.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp, i64, vp]
    L.gg_dropout_bits_words.argtypes = [i64]
    L.gg_dropout_bits_words.restype = i64
    L.gg_dropout_bits.argtypes = [vp, C.c_uint32, f32, i64, vp, vp]
    L.gg_engine_set_batchnorm.argtypes = [vp, vp, vp, C.c_float, C.c_float]
    L.gg_engine_generate_keep.argtypes = [vp, vp, vp, i32, vp]
    L.gg_engine_generate_backward.argtypes = [vp, vp, vp, vp]
    L.gg_engine_critic_keep.argtypes = [vp, vp, vp, i32, vp]
    L.gg_engine_critic_backward.argtypes = [vp, vp, vp, vp]
    L.gg_engine_disc_grads.argtypes = [vp, vp, vp, i32, vp]
    L.gg_engine_gen_grads.argtypes = [vp, vp, i32, vp]
    L.gg_engine_disc_grads_phase.argtypes = [vp, vp, vp, i32, i32, vp]
    L.gg_engine_gen_grads_phase.argtypes = [vp, vp, i32, i32, vp]
    L.gg_engine_optim_step.argtypes = [vp, i32, f32, vp]
    L.gg_engine_generate.argtypes = [vp, vp, vp, i32, vp]
    L.gg_engine_critic.argtypes = [vp, vp, vp, i32, vp]
    L.gg_engine_gradient_penalty.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    L.gg_engine_gp_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.gg_engine_lanes_signal.argtypes = [vp, vp]
    L.gg_masked_mean_rows.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.gg_gather_rows.argtypes = [vp, i64, vp, vp, i64, i64, i32, vp]
    L.gg_engine_stats.argtypes = [vp]
    L.gg_engine_stats.restype = vp
    L.gg_engine_buffer.argtypes = [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.gg_engine_buffer.restype = vp
    L.gg_optim_step.argtypes = [i32, vp, vp, vp, vp, i64, f32, f32, vp, vp, vp, vp]
    L.gg_attention_fwd.argtypes = [C.POINTER(AttnArgs), vp]
    L.gg_attention_bwd.argtypes = [C.POINTER(AttnArgs), vp]
    L.gg_wgrad_group_workspace_bytes.argtypes = [i64]
    L.gg_wgrad_group_workspace_bytes.restype = i64
    L.gg_wgrad_group.argtypes = [C.POINTER(WgradItem), i32, vp, i64, vp]
    L.gg_colsum_group_workspace_bytes.argtypes = [i64]
    L.gg_colsum_group_workspace_bytes.restype = i64
    L.gg_colsum_group.argtypes = [C.POINTER(ColsumItem), i32, vp, i64, vp]
    L.gg_encoder_layer_fwd.argtypes = [C.POINTER(EncLayerParams), vp]
    L.gg_enc_layer_set_trace.argtypes = [vp]
    L.gg_encoder_ffn_bwd.argtypes = [C.POINTER(EncFfnBwdParams), vp]
    L.gg_enc_layer_profile.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_longlong)]
    L.gg_wgrad_group_profile.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_longlong)]
    declare_evalmetrics(L)
    L.gg_launch_count.argtypes = [i32]
    L.gg_launch_count.restype = C.c_longlong
    L.gg_launch_count_add.argtypes = [C.c_longlong]
    L.gg_launch_count_add.restype = None
    L.gg_gemm_profile_begin.argtypes = []
    L.gg_gemm_profile_dump.argtypes = [C.c_char_p]
    L.gg_gemm_set_trace.argtypes = [vp, i32]
    L.gg_gemm_set_timer.argtypes = [vp, i32]
    L.gg_gemm_timer_slots.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), i32]
    L.gg_gemm_profile_bytes.argtypes = []
    L.gg_gemm_profile_bytes.restype = C.c_double
    L.gg_gemm_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]


EXPORTS = [
    "gg_last_error", "gg_abi_version", "gg_check_device", "gg_gemm_bf16", "gg_engine_workspace_bytes",
    "gg_engine_create", "gg_engine_destroy", "gg_engine_set_lanes", "gg_engine_refresh_shadows", "gg_engine_set_batch", "gg_engine_set_labels", "gg_engine_set_batchnorm", "gg_engine_generate_keep", "gg_engine_generate_backward",
    "gg_engine_critic_keep", "gg_engine_critic_backward",
    "gg_engine_disc_grads", "gg_engine_gen_grads", "gg_engine_disc_grads_phase", "gg_engine_gen_grads_phase", "gg_engine_optim_step", "gg_engine_generate",
    "gg_engine_critic", "gg_engine_gradient_penalty", "gg_engine_gp_step", "gg_masked_mean_rows", "gg_gather_rows", "gg_engine_lanes_signal", "gg_engine_stats", "gg_engine_buffer", "gg_optim_step",
    "gg_launch_count", "gg_launch_count_add", "gg_gemm_profile_begin", "gg_gemm_profile_end", "gg_gemm_profile_dump", "gg_gemm_set_trace", "gg_gemm_profile_bytes", "gg_gemm_set_timer", "gg_gemm_timer_slots",
    "gg_pairwise_distance", "gg_row_kth_smallest", "gg_row_membership", "gg_col_hits", "gg_standardize_columns",
    "gg_gene_correlation", "gg_gamma_moments_workspace_bytes", "gg_gamma_moments",
    "gg_encoder_layer_fwd", "gg_encoder_ffn_bwd", "gg_enc_layer_set_trace", "gg_enc_layer_profile", "gg_wgrad_group_profile", "gg_attention_fwd", "gg_attention_bwd", "gg_dropout_bits", "gg_dropout_bits_words", "gg_xw_f32", "gg_gemm_layernorm", "gg_film_patch_encode", "gg_wgrad_group", "gg_wgrad_group_workspace_bytes", "gg_colsum_group", "gg_colsum_group_workspace_bytes",
]
