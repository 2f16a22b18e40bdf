# Convenience targets (the driver uses __graft_entry__.build() / pytest / bench.py directly).
PY ?= python

build:            ## compile libmarlsat_b200.so (nvcc, sm_100a) and the oracle's C restatement (gcc)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle KATs, host logic, ABI symbols, gloo world-size-2 (no GPU needed)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity tests through the C ABI on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench:            ## headline workload, one GPU
	$(PY) bench.py

golden:           ## regenerate tests/golden/ from the reference's own pure-Python checkers (needs /root/reference)
	$(PY) tests/golden/make_golden.py

.PHONY: build test-cpu test-gpu bench golden
