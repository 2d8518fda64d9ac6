"""ctypes binding of libiago_b200.so — signatures mirror include/iago_b200.h one to one."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("IAGO_B200_LIB") or os.path.join(HERE, "libiago_b200.so")   # override: A/B builds of the kernels


class IagoError(RuntimeError):
    pass


class IagoRng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("stream_id", C.c_uint32), ("seed", C.c_uint64), ("game_id0", C.c_uint64),
                ("uniforms", C.c_void_p), ("u_stride", C.c_int64), ("forced", C.c_void_p), ("f_stride", C.c_int64)]


class IagoMctsParams(C.Structure):
    _fields_ = [("lmbda", C.c_double), ("c_puct", C.c_double), ("virtual_loss", C.c_double), ("n_thr", C.c_int32),
                ("leaf_batch", C.c_int32), ("n_playouts", C.c_int32), ("slot_policy", C.c_int32), ("slot_value", C.c_int32),
                ("precision", C.c_int32), ("cache_value", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64),
                ("forced_v", C.c_void_p), ("forced_z", C.c_void_p), ("forced_stride", C.c_int64)]


_P = C.c_void_p
# name -> argtypes (restype is int unless listed in _RESTYPE); this table is also what tests compare with the header
SIGNATURES = {
    "iago_abi_version": [],
    "iago_last_error": [],
    "iago_ctx_create": [C.c_int, C.POINTER(_P)],
    "iago_ctx_destroy": [_P],
    "iago_ctx_sync": [_P],
    "iago_ctx_stream": [_P],
    "iago_load_rollout": [_P, _P, _P],
    "iago_legal_actions": [_P, _P, _P, _P, _P, C.c_int64, _P],
    "iago_place_stone": [_P, _P, _P, _P, _P, C.c_int64, _P],
    "iago_rollout": [_P, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), _P, _P, _P, _P, _P, _P, _P],
    "iago_rollout_host": [_P, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), _P, _P, _P, _P, _P, _P],
    "iago_rollout_host_submit": [_P, C.c_int, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), _P, _P, _P, _P, _P],
    "iago_rollout_host_wait": [_P, C.c_int, _P],
    "iago_rollout_sample": [_P, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), C.c_uint32, _P, _P],
    "iago_rollout_logits": [_P, _P, _P, _P, _P, C.c_int64, _P],
    "iago_load_net": [_P, C.c_int, C.c_int, _P, C.c_int64],
    "iago_policy_forward": [_P, C.c_int, _P, _P, _P, C.c_int64, _P, C.c_int, C.c_int, _P],
    "iago_policy_forward_acts": [_P, C.c_int, _P, _P, _P, C.c_int64, _P, _P, C.c_int, _P],
    "iago_value_forward_acts": [_P, C.c_int, _P, _P, _P, C.c_int64, _P, _P, C.c_int, _P],
    "iago_value_forward": [_P, C.c_int, _P, _P, _P, C.c_int64, _P, C.c_int, _P],
    "iago_selfplay": [_P, C.c_int, C.c_int, C.c_int64, _P, _P, C.c_int, C.c_int, C.POINTER(IagoRng), _P, _P, _P, _P, _P, _P,
                      _P, C.c_int, _P, _P, _P],
    "iago_value_selfplay": [_P, C.c_int, C.c_int, C.c_int64, _P, C.c_int, C.POINTER(IagoRng), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "iago_env_step": [_P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(IagoRng), _P, _P, _P, _P, _P],
    "iago_sample_unmasked": [_P, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), _P, _P, _P, _P],
    "iago_sample_masked": [_P, _P, _P, _P, C.c_int64, C.POINTER(IagoRng), _P, _P, _P],
    "iago_mcts_create": [_P, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(_P)],
    "iago_mcts_destroy": [_P],
    "iago_mcts_set_roots": [_P, _P, _P, _P, C.c_int, _P],
    "iago_mcts_get_roots": [_P, _P, _P, _P, _P, _P],
    "iago_mcts_search": [_P, C.POINTER(IagoMctsParams), _P],
    "iago_mcts_root_stats": [_P, _P, _P, _P, _P],
    "iago_mcts_advance": [_P, _P, _P, _P],
    "iago_mcts_export_tree": [_P, C.c_int, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int32), _P],
    "iago_mcts_overflows": [_P, C.POINTER(C.c_int64)],
    "iago_reinforce_create": [_P, _P, C.c_int64, C.c_int, C.POINTER(_P)],
    "iago_reinforce_destroy": [_P],
    "iago_reinforce_grad": [_P, _P, _P, _P, _P, C.c_int64, _P, C.c_int, _P, _P],
    "iago_reinforce_adam_step": [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P],
    "iago_reinforce_get_state": [_P, _P, _P, _P, C.POINTER(C.c_int64)],
    "iago_reinforce_set_state": [_P, _P, _P, _P, C.c_int64],
    "iago_reinforce_sync_slot": [_P, C.c_int],
    "iago_reinforce_set_option": [_P, C.c_int],
    "iago_reinforce_sync_slot_async": [_P, C.c_int, _P],
    "iago_reinforce_adam_step_dev": [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P],
    "iago_reinforce_openings": [_P, C.c_int64, C.c_uint64, C.c_uint64, _P, _P, _P],
    "iago_reinforce_compact": [_P, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "iago_comm_unique_id": [_P, C.c_int64],
    "iago_comm_create": [_P, _P, C.c_int, C.c_int, C.POINTER(_P)],
    "iago_comm_from_nccl": [_P, _P, C.c_int, C.c_int, C.POINTER(_P)],
    "iago_comm_destroy": [_P],
    "iago_comm_rank": [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "iago_comm_allreduce_sum_f32": [_P, _P, C.c_int64, _P],
    "iago_comm_allreduce_sum_i64": [_P, _P, C.c_int64, _P],
    "iago_trainer_create": [_P, C.c_int, _P, C.c_int64, C.c_int, C.POINTER(_P)],
    "iago_value_grad": [_P, _P, _P, _P, C.c_int64, _P, C.c_int, C.c_double, C.c_uint64, C.c_uint64, _P, _P, _P],
    "iago_policy_eval": [_P, _P, C.c_int, _P, C.c_int64, _P, _P],
    "iago_value_eval": [_P, _P, _P, C.c_int64, _P, _P],
    "iago_rollout_trainer_create": [_P, _P, _P, C.POINTER(_P)],
    "iago_rollout_trainer_destroy": [_P],
    "iago_rollout_trainer_grad": [_P, _P, _P, _P, C.c_int64, _P, C.c_int, _P],
    "iago_rollout_trainer_adam_step": [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P],
    "iago_rollout_trainer_get_state": [_P, _P, C.POINTER(C.c_int64)],
    "iago_rollout_trainer_set_state": [_P, _P, C.c_int64],
    "iago_measure_int_peak": [_P, C.c_int, C.POINTER(C.c_double)],
    "iago_last_kernel_ms": [_P, C.POINTER(C.c_float)],
}
_RESTYPE = {"iago_last_error": C.c_char_p, "iago_ctx_stream": _P}

_lib = None


def load_library():
    """Loads the CUDA library or raises — there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise IagoError(f"{SO_PATH} is missing: build it with `python -m iago_b200.build` "
                        "(iago_b200 has no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    if lib.iago_abi_version() != 1:
        raise IagoError("libiago_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().iago_last_error()
        raise IagoError(f"libiago_b200 error {rc}: {msg.decode() if msg else '?'}")
