# Convenience targets; everything here is a thin wrapper around the commands in README.md.
PY ?= python

.PHONY: build test test-gpu bench bench-reference smoke clean

build:            ## libkltb200.so (nvcc, sm_100a), the C oracle and oracle/_ref
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs golden vectors and vs the live reference, host logic, ABI symbols, gloo sharding
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## on a B200: CUDA path vs oracle / golden vectors through the C ABI
	$(PY) -m pytest tests -q -m gpu

smoke: build      ## one small select + track on cuda:0, checked against the oracle
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## one JSON line (workload B); torchrun for N > 1, see README.md
	$(PY) bench.py

bench-reference:  ## the unmodified reference on the host cores, same metric and config
	$(PY) bench.py --impl reference

clean:
	rm -f pyfeaturetrack_b200/csrc/*.o pyfeaturetrack_b200/libkltb200.so oracle/libkltoracle.so
	@echo "(oracle/_ref is kept: it can only be rebuilt where the reference sources are mounted)"
