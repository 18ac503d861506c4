"""Registry of device program families.

A family = one fused CUDA step kernel (csrc/fam_<name>.cu) plus the host-side description
needed to lower a Python env onto it: its agent classes (kind ids), payload classes (type
ids, in device order), obs/act widths and how to collect family parameters from the Python
objects.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Sequence, Tuple


@dataclasses.dataclass
class FamilyInfo:
    name: str
    family_id: int
    payload_types: Sequence[type]            # index == device payload type id
    obs_dim: int
    act_dim: int
    env_kinds: Tuple[int, ...]               # phx_env_kind values the kernels implement
    collect: Callable                        # (env, agents, spec) -> None; fills params, validates
    trace_capacity: Callable                 # (env, agents) -> int
    # (env, agent, word[, value]) -> column: state access for a family's non-engine kernel
    fast_column: Optional[Callable] = None
    supports_supertypes: bool = False
    # A family that is NOT compiled into libphx.so: path of the user's .cu device program (it
    # ends with PHX_USER_PROGRAM(Prog), csrc/phx_user.cuh).  family_id must be FAMILY_USER; the
    # file is compiled at run time (nvcc -cubin, cached) and loaded with phx_create_user.
    program_source: Optional[str] = None


REGISTRY: Dict[str, FamilyInfo] = {}


def register(info: FamilyInfo) -> FamilyInfo:
    REGISTRY[info.name] = info
    return info


def register_user_family(name: str, program_source: str, payload_types: Sequence[type],
                         obs_dim: int, act_dim: int, env_kinds: Tuple[int, ...] = (0, 1, 2),
                         collect: Optional[Callable] = None,
                         trace_capacity: Optional[Callable] = None) -> FamilyInfo:
    """Register an env class family whose device program is the user's own .cu file -- the
    reference's "subclass Agent and write handlers" (phantom/agents.py:48-60) without rebuilding
    libphx.  Agent classes name the family with `__phx_family__ = name` and their kind id with
    `__phx_kind__`; `payload_types[i]` is the device payload type id i."""
    from . import _lib as L

    return register(FamilyInfo(
        name=name, family_id=L.FAMILY_USER, payload_types=tuple(payload_types), obs_dim=obs_dim,
        act_dim=act_dim, env_kinds=tuple(env_kinds), collect=collect or (lambda env, agents, spec: None),
        trace_capacity=trace_capacity or (lambda env, agents: 8 * len(agents)),
        program_source=program_source))


def get(name: str) -> FamilyInfo:
    if name not in REGISTRY:
        # families register themselves when their module is imported
        import importlib

        importlib.import_module(f"phantom_b200.envs.{name}")
    return REGISTRY[name]
