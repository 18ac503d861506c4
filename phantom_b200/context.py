"""Context (reference: phantom/context.py:11-40)."""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional

from .types import AgentID
from .views import AgentView, EnvView


@dataclasses.dataclass(frozen=True)
class Context:
    agent: Any
    agent_views: Dict[AgentID, Optional[AgentView]]
    env_view: EnvView

    @property
    def neighbour_ids(self) -> List[AgentID]:
        return list(self.agent_views.keys())

    def __getitem__(self, view_id):
        return self.agent_views[view_id]

    def __contains__(self, view_id) -> bool:
        return view_id in self.agent_views
