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


REGISTRY: Dict[str, FamilyInfo] = {}


def register(info: FamilyInfo) -> FamilyInfo:
    REGISTRY[info.name] = info
    return info


def get(name: str) -> FamilyInfo:
    if name not in REGISTRY:
        # families register themselves when their module is imported
        import importlib

        importlib.import_module(f"phantom_b200.envs.{name}")
    return REGISTRY[name]
