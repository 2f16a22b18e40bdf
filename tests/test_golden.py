"""Golden fixtures produced by the reference's own importable pure-Python code
(tests/golden/make_golden.py: planted formulas from src/utils/generate_cnf_dataset.py, expectations
from src/utils/check_sat.py and src/test/verify_solutions.py).  The oracle is checked on CPU; the
CUDA path is checked on the GPU box by driving each env to the fixture's assignment with
multi-flip actions and reading back clause status / num_unsatisfied / solved."""
from pathlib import Path

import numpy as np
import pytest

from oracle.sat_env import SATEnvOracle

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "satcheck_golden.npz")
SHAPES = [(20, 91), (35, 149), (50, 218), (7, 12)]


@pytest.mark.parametrize("n,m", SHAPES)
def test_oracle_clause_evaluation_matches_reference_checkers(n, m):
    cl, a = GOLD[f"clauses_{n}_{m}"], GOLD[f"assign_{n}_{m}"]
    env = SATEnvOracle(n, m, 10)
    status, num_unsat = env.calculate_satisfaction(a.astype(np.int32), cl)
    assert np.array_equal(status, GOLD[f"status_{n}_{m}"])
    assert np.array_equal(num_unsat == 0, GOLD[f"allsat_{n}_{m}"])
    assert np.array_equal(GOLD[f"allsat_{n}_{m}"], GOLD[f"verify_{n}_{m}"])
    assert np.array_equal(num_unsat, (~GOLD[f"status_{n}_{m}"]).sum(1))


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", SHAPES)
def test_cuda_clause_evaluation_matches_reference_checkers(n, m):
    import torch
    import marl_sat_b200 as M
    cl, target = GOLD[f"clauses_{n}_{m}"], GOLD[f"assign_{n}_{m}"].astype(np.int32)
    B = cl.shape[0]
    env = M.SATEnv(n, m, 100, action_mode=1, verbose=False)
    keys = np.arange(2 * B, dtype=np.uint32).reshape(B, 2)
    _, st = env.reset(cl, keys)
    cur = st.variable_assignments.cpu().numpy()
    flip = cur ^ target                                              # multi-flip action = XOR mask (env:245-250)
    av = env.agent_vars.cpu().numpy()
    acts = np.where(av[None] >= 0, flip[:, np.clip(av, 0, n - 1)], 0).astype(np.int32)
    obs, st, rew, done, info = env.step_env(None, st, torch.from_numpy(acts).cuda())
    assert np.array_equal(st.variable_assignments.cpu().numpy(), target)
    assert np.array_equal(st.clauses_satisfied_status.cpu().numpy(), GOLD[f"status_{n}_{m}"])
    assert np.array_equal(info["solved"].cpu().numpy(), GOLD[f"allsat_{n}_{m}"])
    assert np.array_equal(info["num_unsatisfied"].cpu().numpy(), (~GOLD[f"status_{n}_{m}"]).sum(1))
    assert np.array_equal(rew["agent_0"].cpu().numpy(), GOLD[f"allsat_{n}_{m}"].astype(np.float32))
