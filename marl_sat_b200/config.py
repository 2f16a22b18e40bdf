"""Reader for the reference's hydra YAML (``configs/MAPPO_CONFIG.yaml``) without hydra/OmegaConf.

``load_config`` returns the nested dict; ``flatten`` reproduces ``flat_config`` of the runner
(``src/runners/mappo_runner.py:121``: ``{**environment, **network, **training}``) and
``make_env`` builds the env with the runner's argument mapping (runner:124-133).
"""
from __future__ import annotations

from typing import Any, Dict

import yaml

from .env import SATEnv

DEFAULT_CONFIG: Dict[str, Any] = {
    "SEED": 42,
    "environment": {"NUM_VARS": 35, "NUM_CLAUSES": 149, "MAX_STEPS": 512, "VARS_PER_AGENT": 7, "action_mode": 0,
                    "rewards": {"R_CLAUSE": 0.0, "R_SAT": 20.0}},
    "training": {"NUM_ENVS": 128, "NUM_STEPS": 512, "NUM_UPDATES": 300, "UPDATE_EPOCHS": 4, "MINIBATCH_SIZE": 256,
                 "LEARNING_RATE": 1e-4, "GAMMA": 0.995, "GAE_LAMBDA": 0.95, "CLIP_EPS": 0.12, "ENT_COEF": 0.005,
                 "VF_COEF": 0.5, "VF_CLIP": 0.5},
}


def load_config(path: str) -> Dict[str, Any]:
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def flatten(config: Dict[str, Any]) -> Dict[str, Any]:
    return {**config.get("environment", {}), **config.get("network", {}), **config.get("training", {})}


def make_env(config: Dict[str, Any], **kw) -> SATEnv:
    fc = flatten(config)
    rewards = fc.get("rewards", {})
    return SATEnv(num_vars=fc["NUM_VARS"], num_clauses=fc["NUM_CLAUSES"], max_steps=fc["MAX_STEPS"],
                  action_mode=fc.get("action_mode", 0), r_clause=rewards.get("R_CLAUSE", 0.02),
                  r_sat=rewards.get("R_SAT", 1.0), gamma=fc.get("GAMMA", 0.99),
                  vars_per_agent=fc.get("VARS_PER_AGENT"), **kw)
