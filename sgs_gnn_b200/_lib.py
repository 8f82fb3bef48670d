"""ctypes binding of libsgs_b200.so (the C ABI declared in include/sgs_b200.h).

There is no fallback of any kind: if the shared library is missing, `lib()` raises and every
op in this package fails loudly.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C sgs_gnn_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgs_b200.so")

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F32 = C.c_float
U64 = C.c_uint64
SZ = C.c_size_t

# name -> (restype, [argtypes])   -- mirrors include/sgs_b200.h one to one
SIGNATURES = {
    "sgs_last_error": (C.c_char_p, []),
    "sgs_version": (I32, []),
    "sgs_launch_count": (I64, []),
    "sgs_edge_index_split": (I32, [P, I64, I64, P, P, P, P]),
    "sgs_edge_index_gather": (I32, [P, I64, P, I64, P, P, P, P]),
    "sgs_edge_index_check32": (I32, [P, P, I64, I64, P, P]),
    "sgs_edge_gather32": (I32, [P, P, P, I64, P, P, P, P]),
    "sgs_degree_scores": (I32, [P, P, I64, I64, P, P, P]),
    "sgs_csr_workspace_bytes": (SZ, [I64, I64]),
    "sgs_csr_build": (I32, [P, P, I64, I64, P, P, P, P, P, SZ, P]),
    "sgs_csr_build_sorted": (I32, [P, P, I64, I64, P, P, P, P, P, SZ, P]),
    "sgs_keys_unsorted": (I32, [P, I64, P, P]),
    "sgs_gcn_norm": (I32, [P, P, P, P, I64, I64, P, P, P, P, P]),
    "sgs_gcn_norm_apply": (I32, [P, P, P, P, P, I64, I64, P, P]),
    "sgs_spmm": (I32, [P, P, P, P, P, P, P, I64, I64, P, P, I32, F32, U64, P]),
    "sgs_table_f16": (I32, [P, I64, I64, I32, P, P, P]),
    "sgs_spmm_h16": (I32, [P, P, P, P, P, P, P, P, I64, I64, P, P, I32, F32, U64, P]),
    "sgs_spmm_sharded": (I32, [P, P, P, P, P, P, P, P, P, I64, I64, P, P, I32, F32, U64, I64, I64, P, I32, I32, I64, P]),
    "sgs_peer_push_rows": (I32, [P, P, I32, I32, I64, I64, I64, I64, I32, P]),
    "sgs_peer_reduce_rows": (I32, [P, I32, I32, I64, I64, I64, I64, P, P]),
    "sgs_spmm_h16_pair": (I32, [P, P, P, P, P, P, P, P, I64, I64, P, P, P, I32, F32, U64, P]),
    "sgs_gcn_edge_grad_h16": (I32, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, P, P, P, P, I32, P]),
    "sgs_act_bwd": (I32, [P, P, I64, F32, P, P]),
    "sgs_colsum": (I32, [P, I64, I64, P, P]),
    "sgs_gcn_edge_grad": (I32, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, P, P, P, P, I32, P]),
    "sgs_gcn_edge_grad_partial": (I32, [P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, P, P, P, P]),
    "sgs_gcn_edge_grad_partial_h16": (I32, [P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, P, P, P, P]),
    "sgs_gcn_edge_grad_final": (I32, [P, P, P, P, P, P, I64, P, I32, P]),
    "sgs_round_tf32": (I32, [P, I64, P, P]),
    "sgs_gemm": (I32, [P, I64, I64, P, I64, I64, P, I64, I64, I64, I64, I32, I32, P]),
    "sgs_edge_score_workspace_bytes": (SZ, [I64, I64, I64, I32, I32]),
    "sgs_edge_score_fwd": (I32, [P, I64, I64, P, P, P, I64, P, P, P, P, F32, U64, P, P, SZ, I32, P]),
    "sgs_edge_score_bwd": (I32, [P, I64, I64, P, P, P, I64, P, P, P, P, F32, U64, P, P, P, P, P, P, P, P, SZ,
                                 I32, P]),
    "sgs_sum_f32": (I32, [P, I64, P, P, SZ, P]),
    "sgs_softmax_f32": (I32, [P, I64, P, P, SZ, P]),
    "sgs_exponential_f32": (I32, [P, I64, U64, P]),
    "sgs_exponential_ids_f32": (I32, [P, P, I64, U64, P]),
    "sgs_topq_keys": (I32, [P, P, P, I64, F32, F32, I32, P, P, P, P, P]),
    "sgs_topq_find": (I32, [P, P, I64, I32, P]),
    "sgs_topq_hist": (I32, [P, I64, P, P, I32, P]),
    "sgs_topq_workspace_bytes": (SZ, [I64]),
    "sgs_topq_compact": (I32, [P, I64, P, I64, P, I64, P, P, P, SZ, P]),
    "sgs_sample_topq": (I32, [P, P, P, I64, I64, F32, F32, I32, P, P, P, P, P, P, SZ, P]),
    "sgs_gather_selected": (I32, [P, P, P, I64, F32, F32, I32, P, P, P, P]),
    "sgs_scatter_selected": (I32, [P, P, I64, P, P]),
    "sgs_loss_fwd": (I32, [P, I64, I64, P, P, P, P, P, P, I64, I32, P, P]),
    "sgs_loss_finish": (I32, [P, F32, F32, F32, I32, I32, P, P]),
    "sgs_loss_bwd": (I32, [P, I64, I64, P, P, P, P, P, P, I64, I32, P, F32, F32, F32, I32, I32, P, P, P, P]),
    "sgs_loss_fwd_fused": (I32, [P, I64, I64, P, P, P, P, P, P, I64, P, P, P, P, P, P]),
    "sgs_loss_bwd_fused": (I32, [P, I64, I64, P, P, P, I64, P, F32, F32, F32, I32, I32, P, P, P, P, P, P, P]),
}

PREC_FP32, PREC_BF16, PREC_FP16, PREC_TF32 = 0, 1, 2, 3
SAMPLE_TRAIN, SAMPLE_TEST, SAMPLE_RAW = 0, 1, 2
SPMM_RELU, SPMM_DROPOUT, SPMM_ACCUM, SPMM_ADD_ROOT = 1, 2, 4, 8
TOPQ_BINS = 2048

_lock = threading.Lock()
_lib = None


class SgsError(RuntimeError):
    """A libsgs_b200 entry point returned a negative SGS_E_* code."""


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA extension has not been built "
                    "(run `make -C sgs_gnn_b200/csrc`); sgs_gnn_b200 has no CPU fallback")
            h = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)
                fn.restype = res
                fn.argtypes = args
            _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().sgs_last_error()
        raise SgsError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().sgs_launch_count())
