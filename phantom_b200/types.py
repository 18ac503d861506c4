"""Type aliases of the plugin API (reference: phantom/types.py:1-5)."""
from typing import Hashable

AgentID = Hashable
PolicyID = Hashable
StageID = Hashable
