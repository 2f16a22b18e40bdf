"""DIMACS ingest: the committed .cnf fixtures were written by the reference's own generator
(tests/golden/make_golden.py) and parsed there by the reference's verify_solutions.parse_cnf_file;
the parser here must return the same clauses, tolerate SATLIB footers, and stack fixed shapes."""
from pathlib import Path

import numpy as np
import pytest

from marl_sat_b200 import dimacs

GOLD_DIR = Path(__file__).resolve().parent / "golden"
GOLD = np.load(GOLD_DIR / "satcheck_golden.npz")


@pytest.mark.parametrize("n,m", [(20, 91), (35, 149), (50, 218), (7, 12)])
def test_parse_matches_reference_parser(n, m):
    for f in range(2):
        nv, nc, clauses = dimacs.parse_cnf(str(GOLD_DIR / f"ref_uf{n}-{m}_{f}.cnf"))
        assert (nv, nc) == (n, m) and len(clauses) == m
        # formula f occupies rows 8f..8f+7 of the fixture (8 candidate assignments per formula)
        assert np.array_equal(np.array(clauses, np.int32), GOLD[f"clauses_{n}_{m}"][8 * f])
        assert dimacs.parse_cnf(str(GOLD_DIR / f"ref_uf{n}-{m}_{f}.cnf"), strict=True)[2] == clauses


def test_satlib_footer_comments_and_ragged_stacking(tmp_path):
    (tmp_path / "a.cnf").write_text("c comment\np cnf 4 3\n1 -2 3 0\n\n-1 4 0\n2 3 -4 0\n%\n0\n")
    (tmp_path / "b.cnf").write_text("p cnf 4 3\n1 2 0\n-3 0\n-1 -2 -4 0\n")
    (tmp_path / "ignored.txt").write_text("x")
    probs = dimacs.load_cnf_problems(str(tmp_path))
    assert [p["name"] for p in probs] == ["a.cnf", "b.cnf"]
    assert probs[0]["clauses"] == [[1, -2, 3], [-1, 4], [2, 3, -4]]
    arr = dimacs.stack_problems(probs)
    assert arr.dtype == np.int32 and arr.shape == (2, 3, 3)
    assert arr[1].tolist() == [[1, 2, 0], [-3, 0, 0], [-1, -2, -4]]
    with pytest.raises(ValueError):                       # strict = the reference's behaviour: "%" is not an int
        dimacs.parse_cnf(str(tmp_path / "a.cnf"), strict=True)
    (tmp_path / "c.cnf").write_text("p cnf 4 2\n1 2 0\n-3 0\n")
    with pytest.raises(ValueError):
        dimacs.stack_problems(dimacs.load_cnf_problems(str(tmp_path)))


@pytest.mark.parametrize("n,m", [(20, 91), (35, 149), (50, 218), (7, 12)])
def test_native_reader_matches_python_parser(n, m):
    for f in range(2):
        path = str(GOLD_DIR / f"ref_uf{n}-{m}_{f}.cnf")
        nv, nc, clauses = dimacs.parse_cnf(path)
        nv2, nc2, arr = dimacs.parse_cnf_native(path)
        assert (nv2, nc2) == (nv, nc) and arr.dtype == np.int32
        assert np.array_equal(arr, np.array(clauses, np.int32))
        assert np.array_equal(dimacs.parse_cnf_native(path, strict=True)[2], arr)


def test_native_reader_edge_cases(tmp_path):
    (tmp_path / "a.cnf").write_text("c comment\np cnf 4 3\n1 -2 3 0\n\n  -1 4 0  \r\n2 3 -4 0\n%\n0\n")
    nv, nc, arr = dimacs.parse_cnf_native(str(tmp_path / "a.cnf"))
    assert (nv, nc) == (4, 3) and arr.tolist() == [[1, -2, 3], [-1, 4, 0], [2, 3, -4]]
    from marl_sat_b200._lib import MsatError
    with pytest.raises(MsatError):                            # '%' is not an integer (reference behaviour)
        dimacs.parse_cnf_native(str(tmp_path / "a.cnf"), strict=True)
    (tmp_path / "b.cnf").write_text("p cnf 3 2\n1 2 x 0\n")
    with pytest.raises(MsatError):
        dimacs.parse_cnf_native(str(tmp_path / "b.cnf"))
    (tmp_path / "b.cnf").write_text("p cnf 4 3\n1 2 0\n-3 0\n-1 -2 -4 0")    # no trailing newline
    bank = dimacs.load_cnf_bank_array(str(tmp_path))
    assert bank.shape == (2, 3, 3) and bank[1].tolist() == [[1, 2, 0], [-3, 0, 0], [-1, -2, -4]]
    assert np.array_equal(bank, dimacs.stack_problems(dimacs.load_cnf_problems(str(tmp_path))))


def test_torch_generator_distribution():
    import torch
    from marl_sat_b200.synth import mixed_ksat_torch, uniform_ksat_torch
    x = uniform_ksat_torch(64, 30, 50, 3, seed=1)
    assert x.dtype == torch.int32 and x.shape == (64, 50, 3)
    assert int(x.abs().min()) >= 1 and int(x.abs().max()) <= 30
    s, _ = torch.sort(x.abs(), dim=2)
    assert not bool((s[:, :, 1:] == s[:, :, :-1]).any())                  # k distinct variables per clause
    assert torch.equal(x, uniform_ksat_torch(64, 30, 50, 3, seed=1))     # deterministic
    y = mixed_ksat_torch(16, 30, 50, 3, 7, seed=2)
    widths = (y != 0).sum(2)
    assert int(widths.min()) >= 3 and int(widths.max()) <= 7
    assert bool(((y != 0).int().diff(dim=2) <= 0).all())                  # zeros only as right padding
