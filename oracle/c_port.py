"""ctypes front end of the plain-C restatement ``oracle/sat_env_c.c`` (TEST ORACLE / CPU BASELINE ONLY).

``build()`` compiles it with gcc + OpenMP into ``oracle/_build/liboracle_c.so``; ``SATEnvOracleC`` exposes the
batched reset / rollout-step entry points on NumPy arrays.  It is the compiled, multi-threaded CPU baseline
of ``bench.py`` and an independent cross-check of the NumPy oracle (tests/test_oracle_c.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from .sat_env import create_agent_groups

HERE = Path(__file__).resolve().parent
SRC = HERE / "sat_env_c.c"
LIB = HERE / "_build" / "liboracle_c.so"
_lib = None


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB), str(SRC)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + res.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SATEnvOracleC:
    """Batched C restatement with the constructor of the reference SATEnv (env:29-32)."""

    def __init__(self, num_vars, num_clauses, max_steps, vars_per_agent=None, action_mode=0):
        self.lib = load()
        self.n, self.m, self.max_steps, self.action_mode = num_vars, num_clauses, max_steps, action_mode
        groups = create_agent_groups(num_vars, vars_per_agent)
        self.agents = list(groups)
        self.A = len(groups)
        self.V = max(len(v) for v in groups.values())
        self.agent_vars = np.full((self.A, self.V), -1, np.int32)
        self.var2agent = np.full((num_vars,), -1, np.int32)
        for i, vs in enumerate(groups.values()):
            self.agent_vars[i, :len(vs)] = vs
            self.var2agent[vs] = i
        self.D = 2 * num_vars + num_clauses

    def _desc(self, k):
        return (self.n, self.m, k, self.A, self.V, self.max_steps, self.action_mode, _p(self.agent_vars), _p(self.var2agent))

    def _alloc(self, B, k) -> Dict[str, np.ndarray]:
        n, m, A, D = self.n, self.m, self.A, self.D
        return {"assign": np.empty((B, n), np.int32), "status": np.empty((B, m), np.uint8),
                "nunsat": np.empty((B,), np.int32), "step": np.empty((B,), np.int32),
                "done": np.empty((B, A), np.uint8), "clauses": np.empty((B, m, k), np.int32),
                "acm": np.empty((B, A, m), np.int32), "anm": np.empty((B, A, n), np.int32),
                "l2a": np.empty((B, m, k), np.int32), "obs": np.empty((B, A, D), np.int32)}

    def reset(self, clauses: np.ndarray, keys: np.ndarray) -> Dict[str, np.ndarray]:
        clauses = np.ascontiguousarray(clauses, np.int32)
        keys = np.ascontiguousarray(keys, np.uint32)
        B, _, k = clauses.shape
        st = self._alloc(B, k)
        self.lib.oracle_c_reset(*self._desc(k), B, _p(clauses), _p(keys),
                                *[_p(st[x]) for x in ("assign", "status", "nunsat", "step", "done", "clauses", "acm",
                                                      "anm", "l2a", "obs")])
        return st

    def step(self, st: Dict[str, np.ndarray], actions: np.ndarray, problems: Optional[np.ndarray] = None,
             new_idx: Optional[np.ndarray] = None, reset_keys: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        """In-place rollout step (learner:418-464); without ``problems`` it is plain vmapped ``step_env``."""
        B, _, k = st["clauses"].shape
        actions = np.ascontiguousarray(actions, np.int32)
        auto = problems is not None
        out = {"reward": np.empty((B, self.A), np.float32), "done_all": np.empty((B,), np.uint8),
               "solved": np.empty((B,), np.uint8), "num_unsatisfied": np.empty((B,), np.int32),
               "episode_step": np.empty((B,), np.int32)}
        if auto:
            problems = np.ascontiguousarray(problems, np.int32)
            new_idx = np.ascontiguousarray(new_idx, np.int32)
            reset_keys = np.ascontiguousarray(reset_keys, np.uint32)
        self.lib.oracle_c_step(*self._desc(k), B, _p(actions), 1 if auto else 0, 0 if not auto else len(problems),
                               _p(problems), _p(new_idx), _p(reset_keys),
                               *[_p(st[x]) for x in ("assign", "status", "nunsat", "step", "done", "clauses", "acm",
                                                     "anm", "l2a", "obs")],
                               *[_p(out[x]) for x in ("reward", "done_all", "solved", "num_unsatisfied", "episode_step")])
        return out

    def rollout_keys(self, rng: np.ndarray, B: int, P: int):
        rng = np.ascontiguousarray(rng, np.uint32)
        chain = np.empty(10, np.uint32)
        idx = np.empty(B, np.int32)
        keys = np.empty((B, 2), np.uint32)
        self.lib.oracle_c_rollout_keys(_p(rng), B, P, _p(chain), _p(idx), _p(keys))
        return chain, idx, keys

    def num_threads(self) -> int:
        return int(self.lib.oracle_c_num_threads())

    def set_threads(self, n: int) -> None:
        """Override OMP_NUM_THREADS (torchrun pins it to 1 for every rank)."""
        self.lib.oracle_c_set_threads(int(n))
