from . import spaces                                    # noqa: F401
from .multi_agent_env import MultiAgentEnv, State       # noqa: F401
