"""``SATDataWrapper`` drop-in (``/root/reference/src/learners/mappo_gnn_sat_learner.py:93-146``).

The reference wrapper adapts ``SATEnv`` for the learner: it stacks the per-agent action dict into
the array ``step_env`` expects (learner:131) and returns ``((local_obs, global_state), state, ...)``
where ``global_state`` is the ``GNNInput`` built from the state (learner:149-195).  Here
``global_state`` is produced by the feature kernel (``marl_sat_b200.features``) in sparse form.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict

from .env import SATEnv, SATState


@dataclass
class GNNWrapperState:
    """learner:85-90.  ``static_graph`` is the formula bank (the static part of the graph)."""
    env_state: SATState
    static_graph: Any


class JaxMARLWrapper:
    """jaxmarl 0.0.7 ``JaxMARLWrapper``: stores ``_env`` and forwards unknown attributes."""

    def __init__(self, env):
        self._env = env

    def __getattr__(self, name: str):
        return getattr(self._env, name)


class SATDataWrapper(JaxMARLWrapper):
    def __init__(self, env: SATEnv, emit_global_state: bool = True):
        super().__init__(env)
        self._emit_global_state = emit_global_state

    def _global_state(self, state: SATState):
        if not self._emit_global_state:
            return None
        from .features import gnn_input_from_state
        return gnn_input_from_state(state)

    def reset(self, problem_clauses, key):                                   # learner:102-121
        local_obs, env_state = self._env.reset(problem_clauses, key)
        return (local_obs, self._global_state(env_state)), GNNWrapperState(env_state, env_state.bank)

    def step(self, key, state: GNNWrapperState, actions: Dict[str, Any]):    # learner:123-146
        local_obs, nxt, reward, done, info = self._env.step_env(key, state.env_state, actions)
        return (local_obs, self._global_state(nxt)), GNNWrapperState(nxt, state.static_graph), reward, done, info
