"""CPU oracle for the batched SATEnv hot path of kongqg/marl-sat.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the
reference's algorithm (``/root/reference/src/envs/multi_agent_sat_env.py`` and
``/root/reference/src/learners/mappo_gnn_sat_learner.py:383-532``) plus the
JAX 0.4.29 Threefry PRNG it depends on.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
``marl_sat_b200`` never does and fails loudly without its CUDA library.

Parity status
-------------
* PRNG (``oracle.threefry``): pinned against the Random123 Threefry-2x32
  known-answer vectors and the ``jax.random.split`` / ``jax.random.uniform``
  values printed in the JAX documentation (tests/test_threefry_kat.py).
* SATEnv / wrapper / rollout auto-reset / Transition / GAE + normalisation /
  rollout metrics / GNN features / greedy evaluation / BC labels: **pinned
  against outputs of the reference itself**.  ``tests/golden/make_golden_env.py``
  executes the reference's UNMODIFIED source files from ``/root/reference``
  (``multi_agent_sat_env.py``, ``graph_constructor.py``,
  ``mappo_gnn_sat_learner.py`` incl. the closures ``_env_step`` and
  ``_calculate_gae``, ``evaluate_policy``, ``compute_joint_labels_parallel_greedy``)
  on NumPy stand-ins for ``jax`` / ``chex`` / ``jaxmarl`` / ``flax``
  (``tests/ref_shim``; JAX itself is not installable here) and commits the
  results as ``tests/golden/env_*.npz`` for the five BASELINE shapes, both
  action modes and the edge cases (padding quirk, out-of-range actions, steps
  past ``done``).  ``tests/test_golden_env.py``: this oracle reproduces them
  bit for bit (CPU), and so does the CUDA path through the C ABI (``-m gpu``).
  What the stand-ins themselves assume about JAX (x64-disabled promotion,
  gather clamp / negative wrap, ``one_hot`` of -1, Threefry layout) is listed
  in ``tests/ref_shim/README.md``; the PRNG part is KAT-pinned as above.
* Clause evaluation is additionally pinned by the reference's own pure-Python
  checkers ``src/utils/check_sat.py`` / ``src/test/verify_solutions.py``
  (tests/golden/make_golden.py, tests/test_golden.py).
"""
