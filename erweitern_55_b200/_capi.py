"""ctypes binding of libe55.so (include/e55.h).  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

# (E55_LIB: another build of the same library, for A/B measurements of a kernel change within one GPU session)
_LIB_PATH = os.environ.get("E55_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libe55.so")

E55_OK = 0
E55_PENDING = 1
E55_ERR_INVALID, E55_ERR_CUDA, E55_ERR_NO_WEIGHTS, E55_ERR_UNSUPPORTED, E55_ERR_NO_DEVICE = -1, -2, -3, -4, -5
PRECISION_FP32, PRECISION_BF16_TC = 0, 1

_f32p = C.POINTER(C.c_float)
_ctx = C.c_void_p

# name -> (restype, argtypes); every symbol include/e55.h declares
SIGNATURES = {
    "e55_version": (C.c_int, []),
    "e55_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "e55_destroy": (None, [_ctx]),
    "e55_last_error": (C.c_char_p, [_ctx]),
    "e55_set_weights": (C.c_int, [_ctx, C.c_int, C.POINTER(_f32p), C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int)]),
    "e55_set_precision": (C.c_int, [_ctx, C.c_int]),
    "e55_get_precision": (C.c_int, [_ctx, C.POINTER(C.c_int)]),
    "e55_predict": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "e55_host_free": (C.c_int, [C.c_void_p]),
    "e55_set_host_threads": (C.c_int, [_ctx, C.c_int]),
    "e55_host_read_bandwidth": (C.c_int, [_ctx, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "e55_get_host_pipeline": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "e55_get_host_split": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "e55_predict_device": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_predict_packed": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_predict_packed_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_predict_packed_moves": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "e55_service_predict_packed_moves": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "e55_service_submit_packed_moves": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "e55_service_fetch": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "e55_predict_masked": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_predict_masked_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_service_start": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "e55_service_stop": (C.c_int, [_ctx]),
    "e55_service_predict": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_service_predict_packed": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "e55_service_stats": (C.c_int, [_ctx, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "e55_bench_selfplay": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "e55_sync": (C.c_int, [_ctx]),
    "e55_stream": (C.c_int, [_ctx, C.POINTER(C.c_void_p)]),
    "e55_launch_count": (C.c_int, [_ctx, C.POINTER(C.c_int64)]),
    "e55_debug_activations": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "e55_debug_legal_priors": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "e55_set_kernel_timing": (C.c_int, [_ctx, C.c_int]),
    "e55_get_kernel_timing": (C.c_int, [_ctx, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "e55_debug_set_lag": (C.c_int, [_ctx, C.c_int]),
    "e55_debug_profile": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_uint64)]),
    "e55_selftest_umma": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "e55_selftest_slice_book": (C.c_int, [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]),
    "e55_bench_umma": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

class SelfplayStats(C.Structure):
    _fields_ = [("seconds", C.c_double), ("moves", C.c_long), ("evals", C.c_long), ("games", C.c_long),
                ("mean_game_plies", C.c_double), ("mean_gpu_batch", C.c_double),
                ("tree_overflows", C.c_long), ("searches_suspended", C.c_long), ("mate_searches", C.c_long),
                ("mate_exhausted", C.c_long), ("records_truncated", C.c_long)]


GAME_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


class SelfplayConfig(C.Structure):
    _fields_ = [("threads", C.c_int), ("games_per_thread", C.c_int), ("total_moves", C.c_long), ("simulations", C.c_int), ("leaf_batch", C.c_int),
                ("mate_depth", C.c_int), ("reuse_tree", C.c_int), ("use_dirichlet", C.c_int),
                ("sampling_moves", C.c_int), ("max_moves", C.c_int), ("seed", C.c_uint64),
                ("sink", GAME_SINK), ("sink_user", C.c_void_p),
                ("cap_enable", C.c_int), ("cap_full", C.c_int), ("cap_fast", C.c_int), ("cap_frac", C.c_float)]


_pos = C.c_void_p
_mcts = C.c_void_p
SIGNATURES.update({
    "e55_get_arch": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "e55_pos_create": (_pos, []),
    "e55_pos_destroy": (None, [_pos]),
    "e55_pos_copy": (_pos, [_pos]),
    "e55_pos_set_sfen": (C.c_int, [_pos, C.c_char_p]),
    "e55_pos_sfen": (C.c_int, [_pos, C.c_char_p, C.c_int]),
    "e55_pos_print": (C.c_int, [_pos, C.c_char_p, C.c_int]),
    "e55_pos_side_to_move": (C.c_int, [_pos]),
    "e55_pos_ply": (C.c_int, [_pos]),
    "e55_pos_generate_moves": (C.c_int, [_pos, C.c_char_p, C.c_int]),
    "e55_pos_do_move": (C.c_int, [_pos, C.c_char_p]),
    "e55_pos_is_repetition": (C.c_int, [_pos, C.POINTER(C.c_int)]),
    "e55_pos_solve_checkmate_dfs": (C.c_int, [_pos, C.c_int, C.c_char_p, C.c_int]),
    "e55_pos_encode_packed": (C.c_int, [_pos, C.c_void_p, C.c_void_p]),
    "e55_pos_encode_packed_batch": (C.c_int, [C.POINTER(_pos), C.c_int, C.c_void_p, C.c_void_p]),
    "e55_pos_policy_index": (C.c_int, [_pos, C.c_char_p]),
    "e55_mcts_create": (_mcts, [C.c_double]),
    "e55_mcts_destroy": (None, [_mcts]),
    "e55_mcts_clear": (C.c_int, [_mcts]),
    "e55_mcts_configure": (C.c_int, [_mcts, C.c_float, C.c_float, C.c_float, C.c_uint64]),
    "e55_mcts_set_root": (C.c_int, [_mcts, _pos, C.c_int]),
    "e55_mcts_expanded": (C.c_int, [_mcts, C.c_int]),
    "e55_mcts_evaluate": (C.c_int, [_mcts, C.c_int, _pos, C.c_void_p, C.c_float]),
    "e55_mcts_add_noise": (C.c_int, [_mcts, C.c_int]),
    "e55_mcts_select_leaf": (C.c_int, [_mcts, C.c_int, _pos, C.c_int]),
    "e55_mcts_backpropagate": (C.c_int, [_mcts, C.c_int]),
    "e55_mcts_get_usage": (C.c_double, [_mcts]),
    "e55_mcts_get_nodes": (C.c_int, [_mcts]),
    "e55_mcts_get_playouts": (C.c_int, [_mcts, C.c_int, C.c_int]),
    "e55_mcts_best_move": (C.c_int, [_mcts, C.c_int, C.c_char_p, C.c_int]),
    "e55_mcts_softmax_sample": (C.c_int, [_mcts, C.c_int, C.c_float, C.c_char_p, C.c_int]),
    "e55_mcts_softmax_sample_among_top_moves": (C.c_int, [_mcts, C.c_int, C.c_float, C.c_float, C.c_char_p, C.c_int]),
    "e55_mcts_info": (C.c_int, [_mcts, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float)]),
    "e55_mcts_dump": (C.c_int, [_mcts, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "e55_mcts_describe": (C.c_int, [_mcts, C.c_int, C.c_char_p, C.c_int]),
    "e55_mcts_visualize": (C.c_int, [_mcts, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "e55_mcts_search_root": (C.c_int, [_mcts, _ctx, _pos, C.c_int, C.c_int, C.c_int]),
    "e55_mcts_search_batches": (C.c_int, [_mcts, _ctx, _pos, C.c_int, C.c_int, C.c_int, C.c_int]),
    "e55_gpu_selfplay_run": (C.c_int, [_ctx, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(SelfplayStats),
                                       C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "e55_gpu_selfplay_set_cap": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float]),
    "e55_gpu_search_open": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "e55_gpu_search_close": (None, [C.c_void_p]),
    "e55_gpu_search_select": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8),
    "e55_gpu_search_expand": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4),
    "e55_selftest_engine": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "e55_selftest_encode": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "e55_selfplay_default_config": (None, [C.POINTER(SelfplayConfig)]),
    "e55_selfplay_run_ex": (C.c_int, [_ctx, C.POINTER(SelfplayConfig), C.POINTER(SelfplayStats)]),
    "e55_selfplay_run": (C.c_int, [_ctx, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(SelfplayStats)]),
})

_lib = None


class E55Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libe55 error {code}: {message}")
        self.code = code


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libe55.so once.  There is no fallback: a missing library is an ImportError."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `make -C erweitern_55_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc):
    if rc != E55_OK:
        msg = load().e55_last_error(ctx)
        raise E55Error(rc, msg.decode("utf-8", "replace") if msg else "")
