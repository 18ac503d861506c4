"""phantom_b200 -- B200-native batched env-step engine behind Phantom's plugin API.

Public surface mirrors /root/reference/phantom/__init__.py:3-20 for the env-step hot path
(PhantomEnv / Network / Resolver / Agent / FSM / Stackelberg / payload declarations); the
step loop itself runs as hand-written sm_100a CUDA kernels in libphx.so, reached through
the C ABI of include/phx.h.  There is no CPU fallback: importing this package without the
built library raises ImportError.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (fails loudly if libphx.so is missing)
from . import (agents, context, decoders, encoders, errors, fsm, message, metrics, network,
               resolvers, reward_functions, sharding, spaces, views)
from .agents import Agent, StrategicAgent
from .context import Context
from .decoders import Decoder
from .encoders import Encoder
from .env import BatchStep, PhantomEnv
from .env_wrappers import SingleAgentEnvAdapter
from .errors import DeviceOnlyError, NotLowerableError
from .fsm import FiniteStateMachineEnv, FSMStage, StageRule
from .message import Message, MsgPayload, msg_payload
from .network import Network, NetworkError, StochasticNetwork
from .policy import Policy
from .reward_functions import RewardFunction
from .stackelberg import StackelbergEnv
from .supertype import Supertype
from . import utils
from .types import AgentID, PolicyID, StageID
from .views import AgentView, EnvView, View
