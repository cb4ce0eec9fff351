"""All G ranks of the one-process multi-GPU mode on ONE GPU (SMJ_RANKS_ON_ONE_GPU=1): the fabric path's kernels -- samples,
splitters, partition, counts, exchange to G buckets, arrival, local pipelines -- at the C2-per-rank shape, for ncu launch
lists and parity on a single-GPU box.  python tools/dist_onegpu.py [G=8] [rows per rank=10000000] [steps=3] [check=0] [cols=4] [selectivity=0.5] [rows per rank of table 2]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SMJ_RANKS_ON_ONE_GPU"] = "1"
import numpy as np
import smj_b200

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
check = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cols = int(sys.argv[5]) if len(sys.argv) > 5 else 4
sel = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
n2 = int(sys.argv[7]) if len(sys.argv) > 7 else n          # rows per rank of table 2
tot, tot2 = G * n, G * n2
t1 = smj_b200.datagen.table(tot, cols, 1)
t2 = smj_b200.datagen.table(tot2, cols, 2, total_rows=tot)
thr = int(3 * tot * (1.0 - sel))
kn = dict(select_val1=thr, select_val2=thr)
for i in range(steps):
    t0 = time.perf_counter()
    got, st = smj_b200.run(t1, t2, nr_gpus=G, **kn)
    wall = (time.perf_counter() - t0) * 1e3
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k.endswith("_ms") or k.startswith("rows") or k == "kernel_launches"}, "wall_ms", round(wall, 1))
if check:
    from oracle import oracle
    want, sel, _ = oracle.Port().run(t1, t2, 0, kn["select_val1"], 0, kn["select_val2"], 0, 0)
    assert got.shape == want.shape and np.array_equal(got, want)
    print("ONEGPU_OK", G, n, got.shape[0])
smj_b200.lib().smj_shutdown()
