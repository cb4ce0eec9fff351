# Top-level build, with the shape of the reference's sort-merge-join/Makefile:20-41 (`make` builds everything, `make clean`
# removes it).  The reference builds cpu_app, app and four DPU binaries; here `make` builds
#   pim-sort-merge-join_b200/libsmj.so   the engine (nvcc, sm_100a only; replaces select / sort_dpu / merge_dpu / join)
#   host/app                             the C driver (replaces app): host/app data1.csv data2.csv -> ./data/result.csv
#   oracle/_build, oracle/_ref           the CHECKERS (the role of cpu_app; test infrastructure, never linked into the product)
.PHONY: all lib app oracle test bench clean

all: lib app oracle

lib:
	$(MAKE) -s -C pim-sort-merge-join_b200

app: lib
	$(MAKE) -s -C host

oracle:
	$(MAKE) -s -C oracle port ref

test: all
	python -m pytest tests -q -m "not gpu"

bench: all
	python bench.py

clean:
	$(MAKE) -s -C pim-sort-merge-join_b200 clean
	$(MAKE) -s -C host clean
