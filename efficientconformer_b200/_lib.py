"""ctypes binding of libeffconf_b200.so (the C ABI declared in include/effconf_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or the device is not sm_100,
every entry point raises.  Structures mirror the header field-for-field."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeffconf_b200.so")

EC_MAX_BLOCKS = 32
PREC_TF32, PREC_BF16, PREC_BF16X2 = 0, 1, 2
# "bf16x2": split mode, every operand element is a packed (hi, lo) bf16 pair = 16 significant bits (include/effconf_b200.h)
PRECISIONS = {"tf32": PREC_TF32, "bf16": PREC_BF16, "bf16x2": PREC_BF16X2}
_fp = C.POINTER(C.c_float)


class BlockCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim_model", "dim_expand", "num_heads", "kernel_size", "group_size", "conv_stride",
                                          "ff_ratio", "reserved")]


class Config(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("sub_filters", C.c_int32), ("num_blocks", C.c_int32), ("vocab", C.c_int32),
                ("sub_layers", C.c_int32), ("sub_filters2", C.c_int32), ("blocks", BlockCfg * EC_MAX_BLOCKS)]


class FfnRaw(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln_w", "ln_b", "w1", "b1", "w2", "b2")]


class BlockRaw(C.Structure):
    _fields_ = [("ffn1", FfnRaw), ("ffn2", FfnRaw)] + [(n, C.c_void_p) for n in (
        "att_ln_w", "att_ln_b", "u", "v", "wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "wpos", "bpos",
        "conv_ln_w", "conv_ln_b", "pw1_w", "pw1_b", "dw_w", "dw_b", "bn_w", "bn_b", "bn_rm", "bn_rv", "pw2_w", "pw2_b",
        "norm_w", "norm_b", "res_w", "res_b")]


class RawWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("sub_conv_w", "sub_conv_b", "sub_bn_w", "sub_bn_b", "sub_bn_rm", "sub_bn_rv",
                                          "sub2_conv_w", "sub2_conv_b", "sub2_bn_w", "sub2_bn_b", "sub2_bn_rm", "sub2_bn_rv",
                                          "lin_w", "lin_b", "fc_w", "fc_b")] + [("blocks", BlockRaw * EC_MAX_BLOCKS)]


_SIGNATURES = {
    "ec_last_error": (C.c_char_p, []),
    "ec_version": (C.c_int, []),
    "ec_device_check": (C.c_int, []),
    "ec_engine_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    "ec_engine_destroy": (None, [C.c_void_p]),
    "ec_engine_weight_bytes": (C.c_size_t, [C.c_void_p]),
    "ec_engine_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ec_engine_prepare": (C.c_int, [C.c_void_p, C.POINTER(RawWeights), C.c_void_p, C.c_void_p]),
    "ec_engine_relpos_rows": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ec_engine_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_engine_out_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "ec_profile_categories": (C.c_int, []),
    "ec_profile_category_name": (C.c_char_p, [C.c_int]),
    "ec_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ec_engine_last_launches": (C.c_int, [C.c_void_p]),
    "ec_engine_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int32)]),
    "ec_engine_set_fuse_ln": (C.c_int, [C.c_void_p, C.c_int]),
    "ec_engine_set_fuse_ffn": (C.c_int, [C.c_void_p, C.c_int]),
    "ec_engine_set_fuse_front": (C.c_int, [C.c_void_p, C.c_int]),
    "ec_set_pdl": (C.c_int, [C.c_int]),
    "ec_engine_set_skip_mask": (C.c_int, [C.c_void_p, C.c_uint]),
    "ec_debug_gemm_timeline": (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong)]),
    "ec_debug_gemm_block_n": (C.c_int, [C.c_int]),
    "ec_debug_ffn_timeline": (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong)]),
    "ec_op_gemm_ln": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_void_p]),
    "ec_op_ffn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "ec_ctc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ec_ctc_loss": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_layernorm_bwd_work_bytes": (C.c_size_t, [C.c_int]),
    "ec_op_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_layernorm_bwd_emit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_uint, C.c_void_p]),
    "ec_op_cast_scaled": (C.c_int, [C.c_int, C.c_void_p, C.c_float, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ec_op_swish_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ec_op_glu_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_strided_rows": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_strided_rows_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_subsample_conv_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_col_stats_work_bytes": (C.c_size_t, [C.c_int]),
    "ec_op_col_stats": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_group_stats_merge": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ec_op_group_expand": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_group_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_subsample_wgrad_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ec_op_subsample_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "ec_op_relpos_attention_bwd_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ec_op_relpos_attention_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "ec_op_relpos_attention_bwd_act": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
    "ec_op_conv_train_work_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "ec_op_dwconv_raw": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_bn_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "ec_op_bn_swish_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "ec_op_bn_swish_bwd_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_bn_swish_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ec_op_dwconv_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_wgrad_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ec_op_wgrad": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_wgrad_bias": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "ec_op_colsum_work_bytes": (C.c_size_t, [C.c_int]),
    "ec_op_colsum": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_transpose_cast": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_swish_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ec_op_glu_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_ctc_grad_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ec_ctc_loss_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_ctc_greedy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_dropout_advance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ec_op_dropout": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_uint,
                                C.c_void_p]),
    "ec_op_dropout_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_size_t, C.c_void_p, C.c_float, C.c_void_p, C.c_uint,
                                         C.c_void_p]),
    "ec_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "ec_op_stats_merge_ranks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_swish_dropout": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_float, C.c_void_p, C.c_uint, C.c_void_p]),
    "ec_op_transpose_cast_multi": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_pack_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_pack_flat_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ec_op_cast": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ec_weight_planes": (C.c_int, [C.c_int]),
    "ec_op_cast_weight": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ec_op_cast_multi": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_layernorm": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "ec_op_gemm": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_gemm_ex": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ec_op_gemm_train": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_float, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p]),
    "ec_op_gemm_ln_train": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_uint, C.c_void_p]),
    "ec_attention_operand_kind": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ec_op_joint_hidden": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_rnnt_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ec_rnnt_loss": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_rnnt_loss_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ec_op_joint_hidden_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "ec_op_logmel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                               C.c_float, C.c_void_p, C.c_void_p]),
    "ec_op_specaugment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                    C.c_void_p]),
    "ec_p2p_mailbox_bytes": (C.c_size_t, [C.c_int]),
    "ec_p2p_max_payload_floats": (C.c_int, []),
    "ec_p2p_bn_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "ec_p2p_error": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "ec_op_pointwise_glu": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "ec_op_glu_scratch_rows": (C.c_int, [C.c_int]),
    "ec_op_fold_bn": (C.c_int, [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ec_op_relpos_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ec_op_dwconv_bn_swish": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "ec_op_subsample_conv": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load (once) and return the shared library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m efficientconformer_b200.build` "
                               "(or __graft_entry__.build()); the CUDA library is the only implementation of the hot path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    if _profile is not None:
        return _ProfiledLib(_lib, _profile)
    return _lib


class OpProfile:
    """Measurement hook (bench.py, tools/): while active, every operator entry point (ec_op_*, ec_ctc_*, ec_adam_step) is bracketed by
    a CUDA-event pair on the launching stream; `summary()` returns per-entry-point launch counts, device time and the algorithmic
    FLOPs of the tensor-core operators (2*M*N*K from the call's own arguments).  Eager launches only (not under graph capture)."""
    _GEMM_ARGS = {"ec_op_gemm": (3, 4, 5), "ec_op_gemm_ex": (3, 4, 5), "ec_op_gemm_train": (3, 4, 5), "ec_op_gemm_ln_train": (3, 4, 5), "ec_op_wgrad": (3, 4, 5), "ec_op_wgrad_bias": (3, 4, 5), "ec_op_gemm_ln": (3, 4, 5)}

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profile
        _profile = self
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, flops in self.records:
            ent = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0})
            ent["calls"] += 1; ent["ms"] += e0.elapsed_time(e1); ent["flops"] += flops
        return out


class _ProfiledLib:
    def __init__(self, handle, prof):
        self._h, self._p = handle, prof

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not (name.startswith(("ec_op_", "ec_ctc_", "ec_adam")) and not name.endswith(("_bytes", "_rows"))):
            return fn
        prof = self._p

        def call(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args)
            e1.record()
            idx = OpProfile._GEMM_ARGS.get(name)
            flops = 2.0 * args[idx[0]] * args[idx[1]] * args[idx[2]] if idx else 0.0
            prof.records.append((name, e0, e1, flops))
            return r
        return call


_profile = None


def check(status: int):
    if status != 0:
        raise RuntimeError("effconf_b200: " + lib().ec_last_error().decode())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def act_dtype(precision: int):
    """torch dtype that carries one activation element: fp32 words for TF32-rounded values and for the packed (hi, lo) pairs of the
    split mode (view as int32 / use ops.unpack to look at the values), bf16 otherwise."""
    return torch.bfloat16 if precision == PREC_BF16 else torch.float32


def weight_planes(precision: int):
    return 2 if precision == PREC_BF16X2 else 1
