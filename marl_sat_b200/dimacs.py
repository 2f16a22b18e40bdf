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


def parse_cnf_native(file_path: str, strict: bool = False) -> Tuple[int, int, np.ndarray]:
    """``parse_cnf`` through the native reader of the shared library (``msat_dimacs_parse``): returns
    ``(num_vars, num_clauses, clauses int32[rows, max_width])`` with narrower clauses 0-padded."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    with open(file_path, "rb") as f:
        text = f.read()
    nv, nc, rows, width = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(lib.msat_dimacs_parse(text, len(text), 1 if strict else 0, C.byref(nv), C.byref(nc), C.byref(rows),
                                     C.byref(width), None, 0), f"msat_dimacs_parse({file_path})")
    out = np.zeros((rows.value, max(width.value, 1)), dtype=np.int32)
    _lib.check(lib.msat_dimacs_parse(text, len(text), 1 if strict else 0, None, None, C.byref(rows), C.byref(width),
                                     out.ctypes.data_as(C.c_void_p), out.shape[1]), f"msat_dimacs_parse({file_path})")
    return nv.value, nc.value, out[:, :width.value] if width.value else out[:, :0]


def load_cnf_bank_array(cnf_data_dir: str, strict: bool = False) -> np.ndarray:
    """All ``*.cnf`` files of a directory (sorted, like data_parser.py:60) as one ``int32[P, m, k]`` array via the
    native reader; files must share the clause count, narrower files are 0-padded to the widest clause."""
    names = sorted(f for f in os.listdir(cnf_data_dir) if f.endswith(".cnf"))
    arrays = [parse_cnf_native(os.path.join(cnf_data_dir, n), strict=strict)[2] for n in names]
    if not arrays:
        return np.zeros((0, 0, 0), np.int32)
    if len({a.shape[0] for a in arrays}) != 1:
        raise ValueError("problems have different clause counts")
    k = max(a.shape[1] for a in arrays)
    out = np.zeros((len(arrays), arrays[0].shape[0], k), np.int32)
    for i, a in enumerate(arrays):
        out[i, :, :a.shape[1]] = a
    return out


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
