"""ctypes binding of libgct_b200.so (C ABI in include/gct_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, the
product path raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``python gct_plus_b200/csrc/build.py``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCT_B200_LIB") or os.path.join(_HERE, "libgct_b200.so")     # override: A/B of two builds of the library

DTYPE_F32, DTYPE_BF16 = 0, 1
NUM_GLOBAL_SLOTS, ENC_LAYER_SLOTS, DEC_LAYER_SLOTS = 22, 12, 20

vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class GctConfig(C.Structure):
    _fields_ = [("src_vocab", i32), ("trg_vocab", i32), ("n_layers", i32), ("d_model", i32), ("d_ff", i32),
                ("heads", i32), ("latent_dim", i32), ("nconds", i32), ("use_cond2dec", i32), ("use_cond2lat", i32),
                ("dtype", i32), ("pad_id", i32), ("dropout", f32)]


class GctWeights(C.Structure):
    _fields_ = [("params_f32", vp), ("params_bf16", vp), ("grads_f32", vp), ("slot_offsets_host", vp)]


class GctIO(C.Structure):
    _fields_ = [("src", vp), ("trg", vp), ("src_mask", vp), ("trg_mask", vp), ("econds", vp), ("dconds", vp),
                ("eps", vp), ("z_in", vp), ("B", i32), ("S", i32), ("T", i32), ("train", i32), ("seed", C.c_uint32),
                ("run_encoder", i32), ("run_decoder", i32), ("logits", vp), ("mu", vp), ("log_var", vp), ("z", vp),
                ("enc_attn", vp), ("dec_attn1", vp), ("dec_attn2", vp)]


class GctDecode(C.Structure):
    _fields_ = [("B", i32), ("Lz", i32), ("max_len", i32), ("prefix_len", i32), ("greedy", i32), ("eos_id", i32),
                ("seed", C.c_uint32), ("zs", vp), ("src_mask", vp), ("dconds", vp), ("uniforms", vp), ("ys", vp),
                ("status", vp), ("forced", vp), ("probs_out", vp), ("logits_out", vp), ("skip_done", i32), ("n_active", i32),
                ("rowmap", vp)]


class GctBucket(C.Structure):
    _fields_ = [("stage", i32), ("reserved", i32), ("offset", i64), ("count", i64)]


class GctCorpus(C.Structure):
    _fields_ = [("src_ids", vp), ("trg_ids", vp), ("tok_off", vp), ("sca_src_ids", vp), ("sca_trg_ids", vp), ("sca_off", vp),
                ("econds", vp), ("dconds", vp), ("nconds", i32), ("n_rows", i64)]


_PROTOS = {
    "gct_last_error": (C.c_char_p, []),
    "gct_version": (C.c_int, []),
    "gct_sm": (C.c_int, []),
    "gct_num_slots": (C.c_int, [C.c_int]),
    "gct_set_gemm_backend": (C.c_int, [C.c_int]),
    "gct_set_pdl": (C.c_int, [C.c_int]),
    "gct_set_tma_store": (C.c_int, [C.c_int]),
    "gct_set_epilogue_warps16": (C.c_int, [C.c_int]),
    "gct_set_cta_pair_gemm": (C.c_int, [C.c_int]),
    "gct_set_decode_attn_config": (C.c_int, [C.c_int]),
    "gct_set_attention_backend": (C.c_int, [C.c_int]),
    "gct_set_attention_bias_grad_fused": (C.c_int, [C.c_int]),
    "gct_set_zattn_config": (C.c_int, [C.c_int]),
    "gct_set_sm_budget": (C.c_int, [C.c_int]),
    "gct_set_attention_trace": (C.c_int, [C.c_void_p]),
    "gct_set_attention_persistent": (C.c_int, [C.c_int]),
    "gct_set_residual_box": (C.c_int, [C.c_int]),
    "gct_set_latent_cross_attention": (C.c_int, [C.c_int]),
    "gct_set_ffn_saved_activation": (C.c_int, [C.c_int]),
    "gct_set_persistent_gemm": (C.c_int, [C.c_int]),
    "gct_norm_fwd": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "gct_norm_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "gct_gemm": (C.c_int, [vp, C.c_int, i64, vp, C.c_int, i64, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "gct_attention_fwd": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, i64, C.c_int, vp, C.c_int, vp, vp,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "gct_attention_bwd": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, i64, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp,
                                    C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "gct_src_mask": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "gct_trg_mask": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "gct_mask_cast": (C.c_int, [vp, C.c_int, i64, vp, vp]),
    "gct_cast_f32_to_bf16": (C.c_int, [vp, vp, i64, vp]),
    "gct_cast_bf16_to_f32": (C.c_int, [vp, vp, i64, vp]),
    "gct_loss_scratch_bytes": (sz, [i64, i64]),
    "gct_loss_fwd_bwd": (C.c_int, [vp, C.c_int, C.c_int, vp, i64, C.c_int, vp, vp, i64, f32, f32, vp, vp, vp, vp, vp, vp]),
    "gct_prop_head_fwd_bwd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp]),
    "gct_forward_workspace_bytes": (sz, [C.POINTER(GctConfig), C.c_int, C.c_int, C.c_int]),
    "gct_forward": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctWeights), C.POINTER(GctIO), vp, sz, vp]),
    "gct_backward_scratch_bytes": (sz, [C.POINTER(GctConfig), C.c_int, C.c_int, C.c_int]),
    "gct_backward": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctWeights), C.POINTER(GctIO), vp, vp, vp, vp, vp, sz, vp,
                               sz, vp]),
    "gct_adam_step": (C.c_int, [vp, vp, vp, vp, vp, i64, C.c_int, f32, f32, f32, f32, f32, vp]),
    "gct_noam_lr": (C.c_double, [i64, C.c_int, i64]),
    "gct_decode_workspace_bytes": (sz, [C.POINTER(GctConfig), C.c_int, C.c_int, C.c_int]),
    "gct_decode_begin": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctWeights), C.POINTER(GctDecode), vp, sz, vp]),
    "gct_decode_compact": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctDecode), C.c_int, vp, vp, sz, vp]),
    "gct_decode_steps": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctWeights), C.POINTER(GctDecode), C.c_int, C.c_int,
                                   vp, sz, vp]),
    "gct_collate": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                              vp, vp]),
    "gct_detokenize": (i64, [vp, i64, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, i64]),
    "gct_toklen_draw": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, vp, C.c_double, i64, vp]),
    "gct_mt19937_fill": (C.c_int, [vp, vp, vp, vp, i64]),
    "gct_normal_from_mt": (C.c_int, [vp, vp, i64, vp]),
    "gct_decode_launches_per_step": (C.c_int, [C.POINTER(GctConfig)]),
    "gct_decode_launches_per_step_at": (C.c_int, [C.POINTER(GctConfig), C.c_int]),
    "gct_gemm_rownorm": (C.c_int, [vp, i64, vp, i64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, f32, vp]),
    "gct_set_rownorm_fusion": (C.c_int, [C.c_int]),
    "gct_decode_begin_launches": (C.c_int, [C.POINTER(GctConfig), C.c_int]),
    "gct_decode_attention": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, i64, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int,
                                       C.c_int, C.c_int, C.c_int, vp]),
    "gct_allreduce_grads": (C.c_int, [vp, vp, i64, vp]),
    "gct_nccl_unique_id": (C.c_int, [vp]),
    "gct_nccl_comm_init": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, vp]),
    "gct_nccl_comm_destroy": (C.c_int, [vp]),
    "gct_backward_dp": (C.c_int, [C.POINTER(GctConfig), C.POINTER(GctWeights), C.POINTER(GctIO), vp, vp, vp, vp, vp, sz, vp, sz, vp, vp,
                                  C.c_int, vp, vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)
_lib = None


class GctError(RuntimeError):
    pass


def lib():
    """Loads the shared library once; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GctError(f"{LIB_PATH} not found: build the CUDA extension first (__graft_entry__.build()); "
                           "gct_plus_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.gct_sm() != 100:
            raise GctError(f"libgct_b200.so was built for sm_{L.gct_sm()}, expected sm_100a")
        if os.environ.get("GCT_B200_PERSIST") == "0":
            L.gct_set_persistent_gemm(0)
        if os.environ.get("GCT_B200_SIMT_ATTN") == "1":
            L.gct_set_attention_backend(1)
        if os.environ.get("GCT_B200_DA_CFG"):
            L.gct_set_decode_attn_config(int(os.environ["GCT_B200_DA_CFG"]))
        if os.environ.get("GCT_B200_TMA_STORE") == "0":
            L.gct_set_tma_store(0)
        if os.environ.get("GCT_B200_PDL") == "0":
            L.gct_set_pdl(0)
        if os.environ.get("GCT_B200_SIMT_GEMM") == "1":      # test hook: bf16 GEMMs through the SIMT kernel
            L.gct_set_gemm_backend(1)
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise GctError(f"{what or 'gct call'} failed ({rc}): {lib().gct_last_error().decode(errors='replace')}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _p(t):
    """data_ptr or None."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise GctError(f"{name} must be a CUDA tensor: gct_plus_b200 runs on sm_100a only, there is no CPU path")


def host_i64(values) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(values, dtype=np.int64))
