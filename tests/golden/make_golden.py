"""Generates tests/golden/*.npz / *.cnf by EXECUTING the reference's own importable pure-Python code
from /root/reference (read-only).  Run in the builder container only (the GPU box has no
/root/reference); the outputs are committed.

What is pinned (the reference has no fixtures of its own):
  * formulas come from the reference's planted-solution generator
    ``src/utils/generate_cnf_dataset.py::generate_sat_cnf`` (the function is extracted with ``ast``
    because the module writes 1000 files at import time) and are parsed back from DIMACS text with
    the reference's ``src/test/verify_solutions.py::parse_cnf_file``;
  * for each (formula, assignment) pair the reference's two clause checkers give the expected
    results: ``src/utils/check_sat.py::check_satisfiability`` (whole formula, and clause by clause
    for the per-clause status) and ``verify_solutions.verify_solution`` (solution strings).
The oracle (and through it the CUDA path) must reproduce ``solved``, ``clauses_satisfied_status``
and ``num_unsatisfied`` for these assignments.

    python tests/golden/make_golden.py
"""
import ast
import importlib.util
import random
import tempfile
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _extract_function(path, fn_name):
    src = Path(path).read_text(encoding="utf-8")
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fn_name)
    ns = {"random": random, "List": list}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(path), "exec"), ns)
    return ns[fn_name]


def main():
    check_sat = _load(REF / "src/utils/check_sat.py", "ref_check_sat")
    verify = _load(REF / "src/test/verify_solutions.py", "ref_verify_solutions")
    generate_sat_cnf = _extract_function(REF / "src/utils/generate_cnf_dataset.py", "generate_sat_cnf")

    shapes = [(20, 91), (35, 149), (50, 218), (7, 12)]
    rng = np.random.default_rng(20261018)
    out = {}
    for si, (n, m) in enumerate(shapes):
        formulas, assigns, all_sat, status, verify_ok = [], [], [], [], []
        for f in range(6):
            seed = 1000 * si + f
            text = generate_sat_cnf(n, m, clause_size=3, seed=seed)
            if f < 2:
                (OUT / f"ref_uf{n}-{m}_{f}.cnf").write_text(text + "\n")
            with tempfile.NamedTemporaryFile("w", suffix=".cnf", delete=False) as tf:
                tf.write(text + "\n")
            clauses = verify.parse_cnf_file(tf.name)
            Path(tf.name).unlink()
            assert len(clauses) == m and all(len(c) == 3 for c in clauses)
            r = random.Random(seed)                       # replay the generator's hidden solution
            sigma = np.array([r.choice([True, False]) for _ in range(n)], dtype=np.int32)
            cands = [sigma.copy()]
            for flips in (1, 2, 5):
                a = sigma.copy()
                a[rng.choice(n, size=min(flips, n), replace=False)] ^= 1
                cands.append(a)
            cands += [rng.integers(0, 2, size=n).astype(np.int32) for _ in range(4)]
            for a in cands:
                formulas.append(np.array(clauses, dtype=np.int32))
                assigns.append(a)
                all_sat.append(bool(check_sat.check_satisfiability(clauses, a)))
                status.append([bool(check_sat.check_satisfiability([c], a)) for c in clauses])
                ok, _ = verify.verify_solution(clauses, "".join(str(int(x)) for x in a))
                verify_ok.append(bool(ok))
        out[f"clauses_{n}_{m}"] = np.stack(formulas)
        out[f"assign_{n}_{m}"] = np.stack(assigns)
        out[f"allsat_{n}_{m}"] = np.array(all_sat)
        out[f"status_{n}_{m}"] = np.array(status)
        out[f"verify_{n}_{m}"] = np.array(verify_ok)
        assert all_sat[0] and verify_ok[0], "the planted solution must satisfy the formula"
    np.savez_compressed(OUT / "satcheck_golden.npz", **out)
    print("wrote", OUT / "satcheck_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
