"""Golden fixtures for the SATEnv / rollout / GAE / GNN-feature semantics, produced by EXECUTING THE
REFERENCE'S OWN, UNMODIFIED SOURCE from /root/reference (read-only) on top of the NumPy stand-ins for
its third-party imports in tests/ref_shim/ (jax, chex, jaxmarl, flax; see tests/ref_shim/README.md).

Run in the builder container only (the GPU box has no /root/reference); the outputs
``tests/golden/env_*.npz`` are committed and both the oracle (CPU test) and the CUDA path (GPU test)
must reproduce them (tests/test_golden_env.py).

What runs, unmodified:
  * ``src/envs/multi_agent_sat_env.py``           imported as a module: ``SATEnv.__init__/reset/step_env/get_obs``
  * ``src/utils/graph_constructor.py``            imported: ``create_static_graph``
  * ``src/learners/mappo_gnn_sat_learner.py``     imported: ``SATDataWrapper.reset/step``, ``Transition``;
      the nested functions ``_env_step`` (learner:383-480: RNG chain, step, reset-all-then-select,
      Transition) and ``_calculate_gae`` (learner:504-528), the normalisation statements
      (learner:530-532) and the metric block (learner:661-686) are compiled from the module's own AST
      nodes (they are closures of ``make_train_cycle`` and cannot be imported by name);
  * ``src/runners/mappo_runner.py::evaluate_policy`` (runner:30-73) and
    ``src/runners/behavioral_cloning.py::compute_joint_labels_parallel_greedy`` (bc:54-100), compiled
    from their AST nodes (their modules import hydra/omegaconf at top level).
The policy network is replaced by a stub that plays a pre-drawn action table and returns pre-drawn
values (the networks are outside the hot path); everything else is the reference's code.

    python tests/golden/make_golden_env.py
"""
from __future__ import annotations

import ast
import contextlib
import io
import sys
from functools import partial
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(HERE.parent / "ref_shim"))

import jax                      # noqa: E402  (the NumPy stand-in)
import jax.numpy as jnp         # noqa: E402
from src.envs.multi_agent_sat_env import SATEnv                                   # noqa: E402
from src.utils.graph_constructor import create_static_graph                       # noqa: E402
import src.learners.mappo_gnn_sat_learner as learner                              # noqa: E402

assert "numpy-shim" in jax.__version__

LEARNER_SRC = REF / "src/learners/mappo_gnn_sat_learner.py"
RUNNER_SRC = REF / "src/runners/mappo_runner.py"
BC_SRC = REF / "src/runners/behavioral_cloning.py"


# ------------------------------------------------------------------ AST helpers
def _find_def(node, name):
    for n in ast.walk(node):
        if isinstance(n, ast.FunctionDef) and n.name == name:
            return n
    raise KeyError(name)


def _compile_nodes(nodes, filename, ns):
    mod = ast.Module(body=list(nodes), type_ignores=[])
    exec(compile(mod, str(filename), "exec"), ns)
    return ns


def _assigned_name(stmt):
    if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name):
        return stmt.targets[0].id
    return None


_learner_tree = ast.parse(LEARNER_SRC.read_text(encoding="utf-8"))
_train_cycle = _find_def(_find_def(_learner_tree, "make_train_cycle"), "_train_cycle")
_env_step_node = _find_def(_train_cycle, "_env_step")
_gae_node = _find_def(_train_cycle, "_calculate_gae")


def _stmts_between(first_name, last_name):
    names = [_assigned_name(s) for s in _train_cycle.body]
    i0 = names.index(first_name)
    i1 = len(names) - 1 - names[::-1].index(last_name)
    return _train_cycle.body[i0:i1 + 1]


_norm_nodes = _stmts_between("adv_mean", "advantages")[:3]          # learner:530-532
_metric_nodes = _stmts_between("team_rewards", "avg_steps_to_solve")     # learner:664-686
assert [_assigned_name(s) for s in _norm_nodes] == ["adv_mean", "adv_std", "advantages"]


# ------------------------------------------------------------------ the commented-out shaped reward (env:201-223)
def shaped_reward_env_class():
    """``SATEnv`` with the alternative ``_calculate_rewards`` that the reference keeps commented out
    (env:201-223): the comment markers of exactly those lines are stripped and the text is compiled as is."""
    lines = (REF / "src/envs/multi_agent_sat_env.py").read_text(encoding="utf-8").splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith("    # def _calculate_rewards"))
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == "#     return rewards")
    body = []
    for l in lines[i0:i1 + 1]:
        t = l[4:]                                    # drop the method indentation
        assert t.startswith("#"), l
        body.append(t[2:] if t.startswith("# ") else t[1:])
    src = "\n".join(body)
    ns = {"jnp": jnp, "Dict": dict, "SATState": object}
    exec(compile(src, "multi_agent_sat_env.py:201-223 (uncommented)", "exec"), ns)
    return type("SATEnvShaped", (SATEnv,), {"_calculate_rewards": ns["_calculate_rewards"]})


# ------------------------------------------------------------------ policy stub (outside the hot path)
class _StubPi:
    """Stands in for the distrax distribution: ``sample`` plays the pre-drawn action table."""
    def __init__(self, net):
        self.net = net

    def sample(self, seed=None):
        net = self.net
        net.act_keys.append(np.asarray(seed).copy())
        a = net.actions[net.t_sample]
        net.t_sample += 1
        return jnp.asarray(a)

    def log_prob(self, action):
        return jnp.zeros(np.asarray(action).shape, dtype=jnp.float32)


class _StubNetwork:
    """``network.apply(..., method=apply_actor | apply_critic)`` of learner:388-396 without a network."""
    def __init__(self, actions, values, num_envs):
        self.actions, self.values, self.B = actions, values, num_envs
        self.t_sample = 0
        self.critic_calls = 0
        self.act_keys = []

    def apply(self, variables, gnn_input, *rest, method=None):
        if method is learner.GNN_ActorCritic.apply_actor:
            return jnp.zeros((), dtype=jnp.float32)          # placeholder leaf; vmap stacks it
        t, b = divmod(self.critic_calls, self.B)
        self.critic_calls += 1
        return jnp.asarray(self.values[t, b])


class _VmapActorPatch:
    """``jax.vmap(actor_fn, ...)`` returns a stacked placeholder; the reference then calls
    ``pi.sample(seed=act_key)`` / ``pi.log_prob`` on it.  The shimmed ``vmap`` returns an ndarray, so the
    namespace in which ``_env_step`` runs sees a ``jax`` whose ``vmap`` wraps actor outputs in a _StubPi."""
    def __init__(self, net):
        self.net = net
        self.__dict__.update({k: getattr(jax, k) for k in ("random", "lax", "tree_util", "nn", "numpy", "jit")})

    def vmap(self, fn, in_axes=0, out_axes=0):
        inner = jax.vmap(fn, in_axes=in_axes, out_axes=out_axes)
        if isinstance(fn, partial) and fn.keywords.get("method") is learner.GNN_ActorCritic.apply_actor:
            return lambda *a, **k: _StubPi(self.net)
        return inner


# ------------------------------------------------------------------ formulas
def random_ksat(rng, P, n, m, k, kmin=None):
    """Uniform random k-SAT, distinct variables per clause; with ``kmin`` the width of every clause is
    uniform in [kmin, k] and the rest is 0 padding (BASELINE config 5)."""
    out = np.zeros((P, m, k), np.int32)
    for p in range(P):
        for c in range(m):
            w = k if kmin is None else int(rng.integers(kmin, k + 1))
            vs = rng.choice(n, size=w, replace=False)
            sg = rng.integers(0, 2, size=w) * 2 - 1
            out[p, c, :w] = (vs + 1) * sg
    return out


# ------------------------------------------------------------------ one rollout case
def run_rollout_case(name, n, m, k, P, B, T, max_steps, vpa=None, action_mode=0, seed=0, kmin=None,
                     wild_actions=False, gamma=0.995, gae_lambda=0.95, dense_graph=False, shaped=None):
    rng = np.random.default_rng(seed)
    with contextlib.redirect_stdout(io.StringIO()) as printed:
        if shaped is None:
            env = SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=action_mode)
        else:       # (r_clause, r_sat, gamma) of the commented-out shaped reward
            env = shaped_reward_env_class()(n, m, max_steps, vars_per_agent=vpa, action_mode=action_mode,
                                            r_clause=shaped[0], r_sat=shaped[1], gamma=shaped[2])
    wrapper = learner.SATDataWrapper(env)
    A, V = env.num_agents, env.max_vars_per_agent
    clauses = random_ksat(rng, P, n, m, k, kmin)
    problems = {"clauses": jnp.asarray(clauses)}
    if action_mode == 0:
        lo, hi = (-V - 2, V + 4) if wild_actions else (0, V + 1)
        actions = rng.integers(lo, hi, size=(T, B, A)).astype(np.int32)
    else:
        actions = rng.integers(0, 2, size=(T, B, A, V)).astype(np.int32)
    values = rng.standard_normal((T + 1, B)).astype(np.float32)
    config = {"action_mode": action_mode, "NUM_ENVS": B, "NUM_STEPS": T, "GAMMA": gamma, "GAE_LAMBDA": gae_lambda}

    # runner:133-137, 289-295 (initial reset; the same _rng feeds randint and split)
    vmapped_reset = jax.vmap(wrapper.reset, in_axes=(0, 0))
    key0 = jax.random.PRNGKey(seed + 42)
    key, _rng = jax.random.split(key0)
    initial_indices = jax.random.randint(_rng, (B,), 0, P)
    reset_keys = jax.random.split(_rng, B)
    (obs0, gs0), env_state0 = vmapped_reset(problems["clauses"][initial_indices], reset_keys)

    net = _StubNetwork(actions, values, B)
    ns = {"jax": _VmapActorPatch(net), "jnp": jnp, "partial": partial, "network": net, "env": wrapper,
          "config": config, "vmapped_reset_fn": vmapped_reset, "problems": problems,
          "GNN_ActorCritic": learner.GNN_ActorCritic, "Transition": learner.Transition}
    _compile_nodes([_env_step_node, _gae_node], LEARNER_SRC, ns)

    class _TS:
        params = None
    carry = (_TS(), env_state0, obs0, gs0, key)
    (_, final_env_state, final_obs, final_gs, final_rng), traj = jax.lax.scan(ns["_env_step"], carry, None, T)
    last_val = jnp.asarray(values[T])
    assert net.critic_calls == T * B and net.t_sample == T

    advantages, targets = ns["_calculate_gae"](traj, last_val)                    # learner:528
    ns2 = {"jnp": jnp, "advantages": advantages}
    _compile_nodes(_norm_nodes, LEARNER_SRC, ns2)                                 # learner:530-532
    ns3 = {"jnp": jnp, "traj_batch": traj}
    _compile_nodes(_metric_nodes, LEARNER_SRC, ns3)                               # learner:664-686

    agents = env.agents
    stack_obs = lambda d_: np.stack([np.asarray(d_[a]) for a in agents], axis=-2)   # [..., A, D]
    fs = final_env_state.env_state
    s0 = env_state0.env_state
    out = {
        "meta": np.array([n, m, k, P, B, T, max_steps, -1 if vpa is None else vpa, action_mode, A, V], np.int64),
        "gamma_lambda": np.array([gamma, gae_lambda], np.float64),
        "shaped": np.array([0.0, 0.0, 0.0, 0.0] if shaped is None else [1.0, *shaped], np.float64),
        "printed": np.array(printed.getvalue()),
        "agent_vars": np.asarray(env.agent_vars), "action_mask": np.asarray(env.action_mask),
        "variable_to_agent_idx": np.asarray(env.variable_to_agent_idx),
        "clauses": clauses, "key0": np.asarray(key0), "rng_after_init": np.asarray(key),
        "initial_indices": np.asarray(initial_indices), "initial_reset_keys": np.asarray(reset_keys),
        "actions": actions, "values": values,
        "act_keys": np.stack(net.act_keys),
        # initial reset (env:158-181 under vmap)
        "obs0": stack_obs(obs0),
        "s0_variable_assignments": np.asarray(s0.variable_assignments),
        "s0_clauses_satisfied_status": np.asarray(s0.clauses_satisfied_status),
        "s0_num_unsatisfied": np.asarray(s0.num_unsatisfied),
        "s0_agent_clause_masks": np.asarray(s0.agent_clause_masks),
        "s0_agent_neighbor_masks": np.asarray(s0.agent_neighbor_masks),
        "s0_literal_to_agent_idx": np.asarray(s0.literal_to_agent_idx),
        "s0_step": np.asarray(s0.step), "s0_done": np.asarray(s0.done),
        "gs0_static_var_features": np.asarray(gs0.static_var_features),
        "gs0_clause_features": np.asarray(gs0.clause_features),
        # Transition (learner:467-478), stacked [T, B, ...]
        "tr_global_done": np.asarray(traj.global_done), "tr_action": np.asarray(traj.action),
        "tr_value": np.asarray(traj.value), "tr_reward": np.asarray(traj.reward),
        "tr_local_obs": stack_obs(traj.local_obs),
        "tr_gs_assignment": np.asarray(traj.global_state.assignment),
        "tr_gs_clause_features": np.asarray(traj.global_state.clause_features),
        "tr_info_solved": np.asarray(traj.info["solved"]),
        "tr_info_num_unsatisfied": np.asarray(traj.info["num_unsatisfied"]),
        "tr_info_episode_step": np.asarray(traj.info["episode_step"]),
        # final carry (learner:479)
        "final_obs": stack_obs(final_obs), "final_rng": np.asarray(final_rng),
        "final_variable_assignments": np.asarray(fs.variable_assignments),
        "final_clauses_satisfied_status": np.asarray(fs.clauses_satisfied_status),
        "final_num_unsatisfied": np.asarray(fs.num_unsatisfied), "final_step": np.asarray(fs.step),
        "final_done": np.asarray(fs.done), "final_clauses": np.asarray(fs.clauses),
        "final_agent_clause_masks": np.asarray(fs.agent_clause_masks),
        "final_agent_neighbor_masks": np.asarray(fs.agent_neighbor_masks),
        "final_literal_to_agent_idx": np.asarray(fs.literal_to_agent_idx),
        "final_gs_assignment": np.asarray(final_gs.assignment),
        "final_gs_clause_features": np.asarray(final_gs.clause_features),
        "final_gs_static_var_features": np.asarray(final_gs.static_var_features),
        # GAE + normalisation + metrics
        "last_val": np.asarray(last_val), "advantages": np.asarray(advantages), "targets": np.asarray(targets),
        "advantages_normalized": np.asarray(ns2["advantages"]),
        "adv_mean": np.asarray(ns2["adv_mean"]), "adv_std": np.asarray(ns2["adv_std"]),
        "metric_mean_episodic_return": np.asarray(ns3["mean_episodic_return"]),
        "metric_solve_rate": np.asarray(ns3["solve_rate"]),
        "metric_avg_unsatisfied_clauses": np.asarray(ns3["avg_unsatisfied_clauses"]),
        "metric_avg_steps_to_solve": np.asarray(ns3["avg_steps_to_solve"]),
    }
    if dense_graph:
        sg = create_static_graph(num_vars=n, num_clauses=m, clauses=jnp.asarray(clauses[0]))
        out["graph0_A_pos"] = np.asarray(sg.A_pos)
        out["graph0_A_neg"] = np.asarray(sg.A_neg)
    for k_, v in out.items():
        if isinstance(v, np.ndarray) and v.dtype in (np.float64,) and k_ not in ("gamma_lambda", "shaped"):
            raise AssertionError(f"{k_}: the shim produced float64")
    np.savez_compressed(HERE / f"env_{name}.npz", **out)
    nres = int(np.asarray(traj.global_done).sum())
    nsol = int(np.asarray(traj.info["solved"]).sum())
    print(f"env_{name}.npz: A={A} V={V} resets={nres} solved={nsol} "
          f"{(HERE / f'env_{name}.npz').stat().st_size / 1024:.0f} KiB")


# ------------------------------------------------------------------ no-auto-reset stepping (env.step_env only)
def run_stepping_case(name, n, m, k, B, T, max_steps, vpa=None, action_mode=0, seed=0, kmin=None):
    """``env.reset`` then T x ``env.step_env`` with no auto-reset: steps past ``done`` (env:225-284),
    negative / out-of-range mode-0 actions, non-binary mode-1 actions are NOT used (outside the space)."""
    rng = np.random.default_rng(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=action_mode)
    A, V = env.num_agents, env.max_vars_per_agent
    clauses = random_ksat(rng, B, n, m, k, kmin)
    keys = rng.integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)
    if action_mode == 0:
        actions = rng.integers(-V - 3, V + 5, size=(T, B, A)).astype(np.int32)
    else:
        actions = rng.integers(0, 2, size=(T, B, A, V)).astype(np.int32)
    obs, state = jax.vmap(env.reset, in_axes=(0, 0))(jnp.asarray(clauses), jnp.asarray(keys))
    stack_obs = lambda d_: np.stack([np.asarray(d_[a]) for a in env.agents], axis=-2)
    rec = {k_: [] for k_ in ("obs", "assign", "status", "nunsat", "step", "done", "reward", "done_all", "solved",
                            "info_nunsat", "episode_step")}
    for t in range(T):
        obs, state, rewards, dones, infos = jax.vmap(env.step_env)(jnp.zeros((B, 2), dtype=jnp.uint32), state,
                                                                   jnp.asarray(actions[t]))
        rec["obs"].append(stack_obs(obs))
        rec["assign"].append(np.asarray(state.variable_assignments))
        rec["status"].append(np.asarray(state.clauses_satisfied_status))
        rec["nunsat"].append(np.asarray(state.num_unsatisfied))
        rec["step"].append(np.asarray(state.step))
        rec["done"].append(np.asarray(state.done))
        rec["reward"].append(np.stack([np.asarray(rewards[a]) for a in env.agents], axis=-1))
        rec["done_all"].append(np.asarray(dones["__all__"]))
        rec["solved"].append(np.asarray(infos["solved"]))
        rec["info_nunsat"].append(np.asarray(infos["num_unsatisfied"]))
        rec["episode_step"].append(np.asarray(infos["episode_step"]))
    out = {"meta": np.array([n, m, k, B, B, T, max_steps, -1 if vpa is None else vpa, action_mode, A, V], np.int64),
           "clauses": clauses, "keys": keys, "actions": actions}
    out.update({k_: np.stack(v) for k_, v in rec.items()})
    np.savez_compressed(HERE / f"env_{name}.npz", **out)
    print(f"env_{name}.npz: A={A} V={V} done-steps={int(out['done_all'].sum())} "
          f"{(HERE / f'env_{name}.npz').stat().st_size / 1024:.0f} KiB")


# ------------------------------------------------------------------ greedy evaluation + BC labels
def run_eval_and_bc_case(name, n, m, k, num_problems, max_steps, vpa=None, seed=0):
    rng = np.random.default_rng(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=0)
    wrapper = learner.SATDataWrapper(env)
    A, V = env.num_agents, env.max_vars_per_agent
    clauses = random_ksat(rng, num_problems, n, m, k)

    runner_tree = ast.parse(RUNNER_SRC.read_text(encoding="utf-8"))
    eval_node = _find_def(runner_tree, "evaluate_policy")
    eval_node.decorator_list = []                     # @partial(jax.jit, static_argnames=...) is the identity here
    ns = _compile_nodes([eval_node], RUNNER_SRC, {"jax": jax, "jnp": jnp, "partial": partial})

    class _Pi:
        def __init__(self, logits):
            self.logits = logits

    class _EvalNet:
        """Greedy 'policy' with pre-drawn logits: argmax over them is the action (runner:41)."""
        def __init__(self, logits):
            self.logits, self.t = logits, 0

        def apply(self, variables, global_state, agent_vars, action_mask):
            pi = _Pi(jnp.asarray(self.logits[self.t]))
            self.t += 1
            return pi, None

    keys = rng.integers(0, 2 ** 32, size=(num_problems, 2), dtype=np.uint64).astype(np.uint32)
    logits = rng.standard_normal((num_problems, max_steps, A, V + 1)).astype(np.float32)
    ever, steps, sols = [], [], []
    for p in range(num_problems):
        w, s, sol = ns["evaluate_policy"](jnp.asarray(keys[p]), wrapper, _EvalNet(logits[p]), None,
                                          jnp.asarray(clauses[p]), max_steps)
        ever.append(bool(w)); steps.append(int(s)); sols.append(np.asarray(sol))

    bc_tree = ast.parse(BC_SRC.read_text(encoding="utf-8"))
    bc_node = _find_def(bc_tree, "compute_joint_labels_parallel_greedy")
    ns_bc = _compile_nodes([bc_node], BC_SRC, {"np": np, "SATEnv": SATEnv})
    assigns = rng.integers(0, 2, size=(num_problems, n)).astype(np.int32)
    taus = [0.0, -1.5]
    labels = np.stack([np.stack([ns_bc["compute_joint_labels_parallel_greedy"](env, clauses[p], assigns[p], tau)
                                 for p in range(num_problems)]) for tau in taus])
    np.savez_compressed(HERE / f"env_{name}.npz",
                        meta=np.array([n, m, k, num_problems, num_problems, max_steps, max_steps,
                                       -1 if vpa is None else vpa, 0, A, V], np.int64),
                        clauses=clauses, keys=keys, logits=logits, ever_solved=np.array(ever),
                        steps_to_solve=np.array(steps, np.int32), solution=np.stack(sols).astype(np.int32),
                        bc_assignments=assigns, bc_taus=np.array(taus), bc_labels=labels.astype(np.int32))
    print(f"env_{name}.npz: eval solved {sum(ever)}/{num_problems}, steps {steps}")


def main():
    # BASELINE.json configs C1..C5 (shapes; small B / T so the fixtures stay small) + edge cases
    run_rollout_case("c1_uf20_mode0", 20, 91, 3, P=6, B=16, T=24, max_steps=5, seed=1, dense_graph=True)
    run_rollout_case("c1_uf20_mode1", 20, 91, 3, P=6, B=8, T=12, max_steps=4, action_mode=1, seed=2)
    run_rollout_case("loose12_mode0", 12, 20, 3, P=5, B=8, T=60, max_steps=9, seed=3)
    run_rollout_case("loose12_mode1", 12, 20, 3, P=5, B=6, T=30, max_steps=7, action_mode=1, seed=4)
    run_rollout_case("yaml_uf35_vpa7", 35, 149, 3, P=4, B=4, T=8, max_steps=3, vpa=7, seed=5)
    run_rollout_case("c2_uf50", 50, 218, 3, P=3, B=4, T=6, max_steps=3, seed=6)
    run_rollout_case("c3_uf100", 100, 430, 3, P=3, B=3, T=5, max_steps=2, seed=7)
    run_rollout_case("c4_uf250", 250, 1065, 3, P=2, B=2, T=3, max_steps=2, seed=8)
    run_rollout_case("c5_mixedk_vpa7", 100, 430, 7, P=3, B=3, T=5, max_steps=2, vpa=7, kmin=3, seed=9)
    run_rollout_case("c5_mixedk_mode1", 30, 64, 7, P=3, B=4, T=8, max_steps=3, vpa=7, kmin=3, action_mode=1, seed=10)
    run_rollout_case("wild_actions_uneven", 23, 60, 3, P=4, B=6, T=20, max_steps=6, seed=11, wild_actions=True)
    run_rollout_case("pad_quirk_n7", 7, 12, 3, P=4, B=5, T=16, max_steps=4, vpa=4, kmin=2, seed=12, dense_graph=True)
    run_rollout_case("single_agent", 9, 20, 3, P=3, B=4, T=12, max_steps=5, vpa=9, seed=13)
    run_rollout_case("one_var_agents", 6, 14, 3, P=3, B=4, T=12, max_steps=5, vpa=1, seed=14)
    run_rollout_case("shaped_loose12", 12, 20, 3, P=5, B=8, T=40, max_steps=9, seed=15, shaped=(0.02, 1.0, 0.99))
    run_rollout_case("shaped_uf50_mode1", 50, 218, 3, P=3, B=4, T=10, max_steps=4, action_mode=1, seed=16,
                     shaped=(0.05, 20.0, 0.995))
    run_stepping_case("past_done_mode0", 12, 20, 3, B=6, T=14, max_steps=4, seed=21)
    run_stepping_case("past_done_mode1", 23, 70, 5, B=5, T=10, max_steps=3, action_mode=1, kmin=2, seed=22)
    run_eval_and_bc_case("eval_bc_loose", 12, 24, 3, num_problems=8, max_steps=25, seed=31)
    run_eval_and_bc_case("eval_bc_uneven", 23, 60, 3, num_problems=4, max_steps=10, seed=32)


if __name__ == "__main__":
    main()
