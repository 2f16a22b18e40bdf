"""DIMACS CNF ingest (next-tier row).

``parse_cnf`` follows ``/root/reference/src/utils/data_parser.py:8-42``: lines starting with ``c`` are
skipped, the ``p cnf n m`` line gives the header, every other line is a clause whose trailing ``0`` is
dropped.  Beyond the reference (which crashes on them): blank lines and the ``%`` /
``0`` footer of SATLIB ``uf`` files are tolerated when ``strict=False`` (default).
``load_cnf_problems`` mirrors data_parser.py:59-72; ``stack_problems`` builds the fixed-shape
``int32[P, m, k]`` array the runner stacks (runner:118), 0-padding narrower clauses.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np


def parse_cnf(file_path: str, strict: bool = False) -> Tuple[int, int, List[List[int]]]:
    clauses: List[List[int]] = []
    num_vars = num_clauses = 0
    with open(file_path, "r") as f:
        for raw in f:
            line = raw.strip()
            if line.startswith("c"):
                continue
            if line.startswith("p"):
                parts = line.split()
                num_vars, num_clauses = int(parts[2]), int(parts[3])
                continue
            if not strict:
                if not line:
                    continue
                if line.startswith("%"):          # SATLIB footer: "%" then a lone "0"
                    break
            literals = [int(x) for x in line.split()]
            clauses.append(literals[:-1])         # drop the terminating 0 (data_parser.py:39)
    return num_vars, num_clauses, clauses


def load_cnf_problems(cnf_data_dir: str, strict: bool = False) -> List[Dict]:
    names = sorted(f for f in os.listdir(cnf_data_dir) if f.endswith(".cnf"))
    problems = []
    for name in names:
        n, m, clauses = parse_cnf(os.path.join(cnf_data_dir, name), strict=strict)
        problems.append({"name": name, "num_vars": n, "num_clauses": m, "clauses": clauses})
    return problems


def stack_problems(problems: List[Dict]) -> np.ndarray:
    """``int32[P, m, k]``; all problems must share the clause count (runner:118 stacks fixed shapes)."""
    m = {len(p["clauses"]) for p in problems}
    if len(m) != 1:
        raise ValueError(f"problems have different clause counts: {sorted(m)}")
    k = max(len(c) for p in problems for c in p["clauses"])
    out = np.zeros((len(problems), m.pop(), k), dtype=np.int32)
    for i, p in enumerate(problems):
        for j, c in enumerate(p["clauses"]):
            out[i, j, :len(c)] = c
    return out
