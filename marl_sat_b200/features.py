"""GNN-input emitter (next-tier row).

Mirrors what ``SATDataWrapper`` builds for the reference's GNN policy:
``create_static_graph`` (``/root/reference/src/utils/graph_constructor.py:93-114``) and
``_state_to_gnn_input`` / ``_calculate_dynamic_clause_features``
(``/root/reference/src/learners/mappo_gnn_sat_learner.py:149-195``).

B200-first split: everything that depends only on the formula (degrees, adjacency) is emitted once
per *bank* and shared by all envs on that formula; the per-env-step part is 4n + 12m bytes.  Dense
``A_pos`` / ``A_neg`` are optional (``dense_adjacency=True``): a consumer that wants the reference's
exact ``GNNInput`` leaves gathers them by ``problem_idx``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .env import FormulaBank, SATState, _ptr, _stream_ptr


@dataclass
class StaticGraph:
    """Per-formula static graph tensors (``StaticGraphData`` + the static node features)."""
    static_var_features: torch.Tensor            # f32 [P, n, 3]
    A_pos: Optional[torch.Tensor]                # f32 [P, n, m] or None
    A_neg: Optional[torch.Tensor]


@dataclass
class GNNInput:
    """graph_constructor.py:34-41, batched over envs."""
    static_var_features: torch.Tensor            # f32 [B, n, 3]
    assignment: torch.Tensor                     # i32 [B, n]
    clause_features: torch.Tensor                # f32 [B, m, 3]
    A_pos: Optional[torch.Tensor]                # f32 [B, n, m] (only with dense adjacency)
    A_neg: Optional[torch.Tensor]


def static_graph(bank: FormulaBank, dense_adjacency: bool = False) -> StaticGraph:
    cache_key = "_static_graph_dense" if dense_adjacency else "_static_graph"
    if getattr(bank, cache_key, None) is not None:
        return getattr(bank, cache_key)
    lib = _lib.load()
    d, dev, P = bank.plan.dims, bank.data.device, bank.num_problems
    svf = torch.empty((P, d.n, 3), dtype=torch.float32, device=dev)
    a_pos = torch.empty((P, d.n, d.m), dtype=torch.float32, device=dev) if dense_adjacency else None
    a_neg = torch.empty((P, d.n, d.m), dtype=torch.float32, device=dev) if dense_adjacency else None
    _lib.check(lib.msat_gnn_static(bank.plan.handle, _ptr(bank.data), P, _ptr(svf), _ptr(a_pos), _ptr(a_neg),
                                   _stream_ptr(dev)), "msat_gnn_static")
    sg = StaticGraph(svf, a_pos, a_neg)
    setattr(bank, cache_key, sg)
    return sg


def dynamic_features(state: SATState):
    """``(assignment i32[B,n], clause_features f32[B,m,3])`` of a packed state."""
    lib = _lib.load()
    bank, d, dev = state.bank, state.bank.plan.dims, state.packed.device
    B = state.num_envs
    assign = torch.empty((B, d.n), dtype=torch.int32, device=dev)
    cf = torch.empty((B, d.m, 3), dtype=torch.float32, device=dev)
    _lib.check(lib.msat_gnn_dynamic(bank.plan.handle, _ptr(bank.data), bank.num_problems, _ptr(state.packed), B,
                                    _ptr(assign), _ptr(cf), _stream_ptr(dev)), "msat_gnn_dynamic")
    return assign, cf


def gnn_input_from_state(state: SATState, dense_adjacency: bool = False) -> GNNInput:
    """``SATDataWrapper._state_to_gnn_input`` (learner:149-174) for every env of ``state``."""
    sg = static_graph(state.bank, dense_adjacency)
    assign, cf = dynamic_features(state)
    idx = state.problem_idx.reshape(-1).long()
    gi = GNNInput(sg.static_var_features[idx], assign, cf,
                  sg.A_pos[idx] if dense_adjacency else None, sg.A_neg[idx] if dense_adjacency else None)
    if not state.batched:
        gi = GNNInput(*[None if t is None else t[0] for t in
                        (gi.static_var_features, gi.assignment, gi.clause_features, gi.A_pos, gi.A_neg)])
    return gi


def flip_gains(state: SATState, tau: float = 0.0):
    """``(delta_unsat i32[B,n], greedy_labels i32[B,A])``: the change of ``num_unsatisfied`` if each
    variable alone were flipped, and the per-agent greedy action labels of the BC expert
    (``/root/reference/src/runners/behavioral_cloning.py:54-100``; ``tau`` = ``TAU_IMPROVE``)."""
    lib = _lib.load()
    bank, d, dev = state.bank, state.bank.plan.dims, state.packed.device
    B = state.num_envs
    delta = torch.empty((B, d.n), dtype=torch.int32, device=dev)
    labels = torch.empty((B, d.A), dtype=torch.int32, device=dev)
    _lib.check(lib.msat_flip_gains(bank.plan.handle, _ptr(bank.data), bank.num_problems, _ptr(state.packed), B,
                                   float(tau), _ptr(delta), _ptr(labels), _stream_ptr(dev)), "msat_flip_gains")
    if not state.batched:
        return delta[0], labels[0]
    return delta, labels
