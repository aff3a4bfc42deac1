"""ctypes binding of librvs_b200.so (include/rvs_b200.h).  No CPU fallback: a missing library or a
missing GPU raises."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "librvs_b200.so")

MEM_DEVICE, MEM_HOST, MEM_HOST_ASYNC = 0, 1, 2
RULES_REF, RULES_STRICT = 0, 1
PLANES_F32_NCHW, PLANES_BF16_NHWC = 0, 1
EVAL_E0, EVAL_ROLLOUT, EVAL_EXTERNAL, EVAL_NN = 0, 1, 2, 3
MODE_REF, MODE_FAST = 0, 1
OPT_LANES_PER_GAME, OPT_NET_GRAPH, OPT_SEARCH_MODE, OPT_GAME_LIMIT, OPT_NET_MAX_CTAS, OPT_NET_PIPELINE, OPT_NET_TOWER = 1, 2, 3, 4, 5, 6, 7
FLAG_OVER, FLAG_WINNER_SHIFT, FLAG_WINNER_MASK, FLAG_PASSED = 1, 1, 6, 8

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class EngineConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("n_games", C.c_int32),
                ("max_sims", C.c_int32), ("max_wave", C.c_int32), ("rules", C.c_int32),
                ("evaluator", C.c_int32), ("c_puct", C.c_float), ("seed", C.c_uint64),
                ("nodes_per_game", C.c_int32), ("net_blocks", C.c_int32), ("net_filters", C.c_int32),
                ("sample_capacity", C.c_int32)]


class EngineStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("sims", "evals", "board_steps", "nodes", "games_finished",
                                          "samples", "launches", "overflow", "tree_bytes", "samples_dropped", "stalled", "nn_evals",
                                          "bad_positions")]


# every symbol include/rvs_b200.h declares: name -> (restype, argtypes)
V = C.c_void_p
PROTOTYPES = {
    "rvs_last_error": (C.c_char_p, []),
    "rvs_version": (C.c_int, []),
    "rvs_launch_count": (C.c_int64, []),
    "rvs_legal_masks": (C.c_int, [V, V, V, V, C.c_int64, C.c_int, C.c_int, V]),
    "rvs_flip_masks": (C.c_int, [V, V, V, V, V, C.c_int64, C.c_int, C.c_int, V]),
    "rvs_apply_moves": (C.c_int, [V, V, V, V, V, V, V, C.c_int64, C.c_int, C.c_int, V]),
    "rvs_random_playouts": (C.c_int, [C.c_int64, C.c_uint64, C.c_uint64, C.c_int, V, V, V, V,
                                      C.POINTER(C.c_int64), C.c_int, V]),
    "rvs_perft": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), V]),
    "rvs_encode_planes": (C.c_int, [V, V, V, V, C.c_int64, C.c_int, C.c_int, C.c_int, V]),
    "rvs_engine_create": (C.c_int, [C.POINTER(EngineConfig), C.POINTER(V)]),
    "rvs_engine_destroy": (C.c_int, [V]),
    "rvs_engine_reset": (C.c_int, [V, V]),
    "rvs_engine_set_positions": (C.c_int, [V, V, V, V, C.c_int32, C.c_int, V]),
    "rvs_engine_get_positions": (C.c_int, [V, V, V, V, V, C.c_int32, C.c_int, V]),
    "rvs_engine_search": (C.c_int, [V, C.c_int32, C.c_int32, V]),
    "rvs_engine_begin_search": (C.c_int, [V, V]),
    "rvs_engine_select": (C.c_int, [V, C.c_int32, V]),
    "rvs_engine_leaf_planes": (C.c_int, [V, V, V, C.c_int, V]),
    "rvs_engine_process": (C.c_int, [V, V, V, C.c_int, V]),
    "rvs_engine_root_visits": (C.c_int, [V, V, C.c_int32, C.c_int, V]),
    "rvs_engine_play": (C.c_int, [V, C.c_float, C.c_int, V, C.c_int, V]),
    "rvs_engine_selfplay": (C.c_int, [V, C.c_int32, C.c_float, C.c_int64, C.c_int, V]),
    "rvs_engine_drain_samples": (C.c_int, [V, V, V, V, C.c_int64, C.POINTER(C.c_int64), C.c_int, V]),
    "rvs_engine_drain_packed": (C.c_int, [V, V, V, V, V, V, C.c_int64, C.POINTER(C.c_int64), C.c_int, V]),
    "rvs_engine_drain_packed_async": (C.c_int, [V, V, V, V, V, V, C.c_int64, V, V]),
    "rvs_engine_set_option": (C.c_int, [V, C.c_int32, C.c_int64]),
    "rvs_engine_predict_probs": (C.c_int, [V, V, V, V, C.c_int64, V, V, C.c_int, V]),
    "rvs_engine_stats_get": (C.c_int, [V, C.POINTER(EngineStats), V]),
    "rvs_engine_set_root_noise": (C.c_int, [V, C.c_double, C.c_float]),
    "rvs_engine_set_lanes_per_game": (C.c_int, [V, C.c_int32]),
    "rvs_engine_load_weights": (C.c_int, [V, V, C.c_int64, C.c_int, V]),
    "rvs_engine_predict": (C.c_int, [V, V, V, V, C.c_int64, V, V, C.c_int, V]),
}

_lib = None


class RvsError(RuntimeError):
    pass


def lib():
    """Loads the CUDA library.  Raises if it was not built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RvsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RvsError(f"librvs_b200 error {rc}: {lib().rvs_last_error().decode(errors='replace')}")


_TORCH_NAMES = {"uint64": ("int64", "uint64"), "uint8": ("uint8",), "int8": ("int8",), "int32": ("int32",),
                "float32": ("float32",), "int64": ("int64",)}


def ptr(x, dtype=None):
    """(pointer, mem kind) of a numpy array or a torch tensor; None -> (None, None).
    `dtype` (numpy dtype name) is what the C side will read: a mismatch raises instead of being
    reinterpreted silently (torch has no uint64 arithmetic, so int64 tensors stand in for uint64)."""
    if x is None:
        return None, None
    if hasattr(x, "data_ptr"):  # torch tensor
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dtype is not None and str(x.dtype).replace("torch.", "") not in _TORCH_NAMES[dtype]:
            raise TypeError(f"expected a {dtype} tensor, got {x.dtype}")
        return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    if dtype is not None and x.dtype.name != dtype:
        raise TypeError(f"expected a {dtype} array, got {x.dtype.name}")
    return x.ctypes.data, MEM_HOST


def mem_of(*xs):
    kinds = {ptr(x)[1] for x in xs if x is not None}
    if len(kinds) != 1:
        raise ValueError("all bulk arguments of one call must live on the same side (host or device)")
    return kinds.pop()


def current_stream(device=None):
    """torch's current CUDA stream handle ON `device` when torch is loaded, else the default stream"""
    import sys
    t = sys.modules.get("torch")
    if t is not None and t.cuda.is_available():
        return t.cuda.current_stream(device).cuda_stream
    return None


def current_device():
    """the CUDA device new engines are created on: torch's current device when torch is loaded (a torchrun
    rank that called torch.cuda.set_device(local_rank) gets ITS GPU), else device 0"""
    import sys
    t = sys.modules.get("torch")
    if t is not None and t.cuda.is_available():
        return t.cuda.current_device()
    return 0
