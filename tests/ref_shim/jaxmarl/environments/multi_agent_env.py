"""``jaxmarl.environments.multi_agent_env`` stand-in: ``State`` (done, step) and the base class
constructor (the reference overrides reset / step_env / get_obs and calls ``step_env`` directly)."""
from _dataclass import dataclass


@dataclass
class State:
    done: object
    step: object


class MultiAgentEnv:
    def __init__(self, num_agents):
        self.num_agents = num_agents
        self.observation_spaces = dict()
        self.action_spaces = dict()
