"""Views (reference: phantom/views.py:5-34).  On the device a view is the start-of-step
snapshot of an agent's public state columns; these classes keep user type annotations and
`EnvView` construction working."""
from __future__ import annotations

import dataclasses
from abc import ABC


@dataclasses.dataclass(frozen=True)
class View(ABC):
    pass


@dataclasses.dataclass(frozen=True)
class AgentView(View):
    pass


@dataclasses.dataclass(frozen=True)
class EnvView(View):
    current_step: int
    proportion_time_elapsed: float
