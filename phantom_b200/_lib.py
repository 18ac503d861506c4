"""ctypes binding of libphx.so (include/phx.h).

The library is the product: there is NO CPU fallback.  If libphx.so has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C phantom_b200/csrc`) the
import of this module fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHX_LIB overrides the library path (A/B builds of the same sources; tools/ab_variants.sh)
LIB_PATH = os.environ.get("PHX_LIB") or os.path.join(_HERE, "libphx.so")

PHX_MAX_AGENTS = 128
PHX_MAX_TYPES = 16
PHX_MAX_STAGES = 8
PHX_MASK_WORDS = 4
PHX_MAX_PARAMS = 16
PHX_TRACE_WORDS = 4
PHX_MAX_CODEC_OPS = 6
PHX_MAX_BASE_CONNECTIONS = 528
PHX_ABI_VERSION = 7

# phx_status
PHX_OK, PHX_ERR_INVALID, PHX_ERR_CUDA, PHX_ERR_UNSUPPORTED, PHX_ERR_NO_DEVICE = 0, -1, -2, -3, -4
# phx_fault
(FAULT_NONE, FAULT_NO_EDGE, FAULT_BAD_PAYLOAD_TYPE, FAULT_UNKNOWN_MSG_TYPE, FAULT_ROUND_LIMIT,
 FAULT_BAD_TRANSITION, FAULT_QUEUE_OVERFLOW, FAULT_INVALID_ACTION, FAULT_UNRESOLVED_MAIL,
 FAULT_PLAN_MISMATCH) = range(10)
# phx_rule_lhs / phx_cmp (device form of an FSM stage handler)
RULE_ALWAYS, RULE_STEP, RULE_AGENT_WORD, RULE_ENV_WORD, RULE_CONST = range(5)
PHX_RULE_BRANCHES, PHX_RULE_TERMS = 4, 2
CMP_LT, CMP_LE, CMP_EQ, CMP_NE, CMP_GE, CMP_GT = range(6)
CMP_F32 = 8  # flag: the term compares float32 values
# phx_env_kind
ENV_BASE, ENV_FSM, ENV_STACKELBERG = 0, 1, 2
# phx_family
FAMILY_SUPPLY_CHAIN, FAMILY_MOCK, FAMILY_MARKET, FAMILY_STACKELBERG, FAMILY_DENSE = 1, 2, 3, 4, 5
FAMILY_SUPPLY_CHAIN2 = 6
FAMILY_SIMPLE_MARKET = 7
FAMILY_DIGITAL_ADS = 8
FAMILY_USER = 100
# phx_exec_mode
EXEC_AUTO, EXEC_QUEUE, EXEC_FAST, EXEC_THREAD, EXEC_WIDE = 0, 1, 2, 3, 4
EXEC_MODES = {"auto": EXEC_AUTO, "queue": EXEC_QUEUE, "fast": EXEC_FAST, "thread": EXEC_THREAD,
              "wide": EXEC_WIDE}
# flags
FLAG_IGNORE_CONNECTION_ERRORS, FLAG_NO_PAYLOAD_CHECKS, FLAG_TRACK_MESSAGES, FLAG_AUTO_RESET = 1, 2, 4, 8
FLAG_STOCHASTIC_NETWORK, FLAG_SHUFFLE_BATCHES = 16, 32
# phx_field
FIELD_STEP, FIELD_EPISODE, FIELD_STAGE, FIELD_TERMINATED, FIELD_TRUNCATED, FIELD_ERROR = range(6)
FIELD_ADJACENCY = 6
FIELD_ENV_STATE = 7
FIELD_FAMILY = 16

_MaskWords = C.c_uint32 * PHX_MASK_WORDS


class PhxRuleTerm(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("lhs", "slot", "word", "cmp", "rhs_kind", "rhs_slot", "rhs_word", "rhs")]


class PhxRuleBranch(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("then", C.c_int32), ("term", PhxRuleTerm * 2)]


class PhxStage(C.Structure):
    _fields_ = [
        ("acting", _MaskWords),
        ("rewarded", _MaskWords),
        ("rewarded_is_none", C.c_int32),
        ("next_stage", C.c_int32),
        ("handler", C.c_int32),
        ("next_allowed", C.c_uint32),
        ("rule_resolves", C.c_int32),
        ("rule_lhs", C.c_int32),
        ("rule_slot", C.c_int32),
        ("rule_word", C.c_int32),
        ("rule_cmp", C.c_int32),
        ("rule_rhs", C.c_int32),
        ("rule_then", C.c_int32),
        ("rule_else", C.c_int32),
        ("rule_n_branches", C.c_int32),
        ("rule_branch", PhxRuleBranch * 4),
        ("n_act_order", C.c_int32),
        ("act_order", C.c_uint8 * PHX_MAX_AGENTS),
    ]


class PhxSpec(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("family", C.c_int32),
        ("env_kind", C.c_int32),
        ("exec_mode", C.c_int32),
        ("flags", C.c_uint32),
        ("num_steps", C.c_int32),
        ("round_limit", C.c_int32),
        ("trace_capacity", C.c_int32),
        ("n_agents", C.c_int32),
        ("n_strategic", C.c_int32),
        ("agent_kind", C.c_int32 * PHX_MAX_AGENTS),
        ("strategic_index", C.c_int32 * PHX_MAX_AGENTS),
        ("adjacency", _MaskWords * PHX_MAX_AGENTS),
        ("n_payload_types", C.c_int32),
        ("type_sender_ok", _MaskWords * PHX_MAX_TYPES),
        ("type_receiver_ok", _MaskWords * PHX_MAX_TYPES),
        ("n_stages", C.c_int32),
        ("initial_stage", C.c_int32),
        ("stages", PhxStage * PHX_MAX_STAGES),
        ("leaders", _MaskWords),
        ("followers", _MaskWords),
        ("obs_dim", C.c_int32),
        ("act_dim", C.c_int32),
        ("iparams", C.c_int32 * PHX_MAX_PARAMS),
        ("fparams", C.c_double * PHX_MAX_PARAMS),
        ("agent_iparam", (C.c_int32 * 4) * PHX_MAX_AGENTS),
        ("agent_fparam", (C.c_double * 4) * PHX_MAX_AGENTS),
        ("agent_codec_op", (C.c_int32 * PHX_MAX_CODEC_OPS) * PHX_MAX_AGENTS),
        ("agent_codec_val", (C.c_float * PHX_MAX_CODEC_OPS) * PHX_MAX_AGENTS),
        ("n_base_connections", C.c_int32),
        ("base_u", C.c_uint8 * PHX_MAX_BASE_CONNECTIONS),
        ("base_v", C.c_uint8 * PHX_MAX_BASE_CONNECTIONS),
        ("base_rate", C.c_double * PHX_MAX_BASE_CONNECTIONS),
    ]


# every symbol include/phx.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "phx_abi_version": (C.c_int32, []),
    "phx_sizeof_spec": (C.c_uint32, []),
    "phx_last_error": (C.c_char_p, []),
    "phx_device_count": (C.c_int32, []),
    "phx_create": (C.c_int32, [C.POINTER(PhxSpec), C.c_int32, C.c_int32, C.c_uint64, C.c_int64,
                               C.POINTER(_P)]),
    "phx_create_user": (C.c_int32, [C.POINTER(PhxSpec), C.c_char_p, C.c_int32, C.c_int32, C.c_uint64,
                                    C.c_int64, C.POINTER(_P)]),
    "phx_destroy": (None, [_P]),
    "phx_num_envs": (C.c_int32, [_P]),
    "phx_exec_name": (C.c_char_p, [_P]),
    "phx_reset": (C.c_int32, [_P, _P, _P, _P, _P]),
    "phx_step": (C.c_int32, [_P] + [_P] * 9 + [_P]),
    "phx_rollout": (C.c_int32, [_P, C.c_int32] + [_P] * 9 + [_P]),
    "phx_rollout_host": (C.c_int32, [_P, C.c_int32] + [_P] * 9),
    "phx_host_alloc": (_P, [C.c_uint64]),
    "phx_host_free": (None, [_P]),
    "phx_get_field": (C.c_int32, [_P, C.c_int32, C.c_int32, _P, C.c_uint64]),
    "phx_set_field": (C.c_int32, [_P, C.c_int32, C.c_int32, _P, C.c_uint64]),
    "phx_get_trace": (C.c_int32, [_P, C.c_int32, C.c_int32, _P, _P]),
    "phx_get_trace_step": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "phx_trace_steps": (C.c_int32, [_P]),
    "phx_jit_source": (C.c_int32, [_P, _P, C.c_uint64, _P]),
    "phx_load_specialised": (C.c_int32, [_P, C.c_char_p]),
    "phx_reduce_field": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "phx_poll_errors": (C.c_int32, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.c_int32]),
    "phx_selftest_ratio": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "phx_selftest_jit_source": (C.c_int32, [C.POINTER(PhxSpec), C.c_int32, C.c_uint64, _P, C.c_uint64, _P]),
    "phx_selftest_wire_expand": (C.c_int32, [C.c_int32, C.c_int32, _P, C.c_uint64, C.c_int32,
                                             _P, _P, _P]),
    "phx_selftest_wire_pack": (C.c_uint32, [C.c_int32] * 5),
}


class LibraryMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: phantom_b200 has no CPU fallback. Build the CUDA "
            "extension first (python -c 'import __graft_entry__ as g; g.build()')."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    if lib.phx_sizeof_spec() != C.sizeof(PhxSpec):
        raise ImportError(
            f"phx_spec layout skew: library {lib.phx_sizeof_spec()} B, binding {C.sizeof(PhxSpec)} B"
        )
    return lib


lib = _load()


class PhxError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libphx status {status}: {message}")
        self.status = status


def check(status: int) -> None:
    if status != PHX_OK:
        raise PhxError(status, (lib.phx_last_error() or b"").decode())


def set_mask(words, index: int) -> None:
    words[index >> 5] |= 1 << (index & 31)
