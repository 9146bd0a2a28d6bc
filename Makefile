# `make test` is the reference's whole build/test driver (`mpirun -n 2 py.test -s`, /root/reference/Makefile:1-2).
# Here the multi-rank tests spawn their own ranks, so plain pytest is enough.
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build
	$(PY) -m pytest tests/ -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests/ -x -q -m gpu

# the reference's literal recipe, for a 2-rank SPMD run of any script
mpirun2:
	$(PY) -m pytorch_ps_mpi_b200.launch -n 2 $(SCRIPT)

bench:
	$(PY) bench.py

.PHONY: build test test-gpu mpirun2 bench
