"""The N>1 path on CPU: two gloo ranks run the key-range exchange PROTOCOL of csrc/smj_dist.cu with the product's
own host-side planning code (smj_plan_splitters / smj_plan_exchange from libsmj.so).  The per-rank device stages
(select, sort, merge, join) are stood in for by the oracle, because there is no GPU here; what is under test is the
planning logic and that shards concatenated in rank order equal the single-process result.  The real NCCL path is
covered by tests/test_multi_gpu.py on >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, pickle
import numpy as np
sys.path.insert(0, os.environ["SMJ_ROOT"])
import torch, torch.distributed as dist
import smj_b200
from oracle import oracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
case = os.environ["SMJ_CASE"]
rng = np.random.default_rng(11)
n1, n2, c1, c2 = 40_000, 30_000, 3, 4
if case == "unique":
    t1 = smj_b200.datagen.table(n1, c1, 1); t2 = smj_b200.datagen.table(n2, c2, 2, total_rows=n1)
    knobs = dict(sel_col1=0, sel_val1=n1, sel_col2=0, sel_val2=n1 // 2, key1=0, key2=0)
elif case == "dups":      # heavy duplicates: runs of one key must not straddle ranks
    t1 = rng.integers(-20, 60, size=(n1, c1)).astype(np.int32); t2 = rng.integers(-20, 60, size=(n2, c2)).astype(np.int32)
    knobs = dict(sel_col1=1, sel_val1=-5, sel_col2=2, sel_val2=0, key1=2, key2=1)
else:                     # one rank selects nothing
    t1 = smj_b200.datagen.table(n1, c1, 1); t1[: n1 // 2, 0] = -1; t2 = smj_b200.datagen.table(n2, c2, 2, total_rows=n1)
    knobs = dict(sel_col1=0, sel_val1=0, sel_col2=0, sel_val2=0, key1=0, key2=0)
port = oracle.Port()
tabs, sel, key = [t1, t2], [(knobs["sel_col1"], knobs["sel_val1"]), (knobs["sel_col2"], knobs["sel_val2"])], [knobs["key1"], knobs["key2"]]

S = 64
recv = []
for t in range(2):
    n = tabs[t].shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world          # this rank's contiguous row block
    loc = port.sort(port.select(tabs[t][lo:hi], *sel[t]), key[t])   # stand-in for the device select+sort+gather
    fk = (loc[:, key[t]].astype(np.int64) + 2**31).astype(np.uint32)  # == (uint32)key ^ 0x80000000
    m = len(fk)
    samp = np.full(S, 0xffffffff, np.uint32)
    if m:
        pos = np.minimum(((2 * np.arange(S) + 1) * m) // (2 * S), m - 1)
        samp = fk[pos]
    recv.append((loc, fk, samp))
# splitters from both tables' samples of every rank (ncclAllGather in the product)
mine = torch.from_numpy(np.concatenate([recv[0][2], recv[1][2]]).astype(np.int64))
allg = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allg, mine)
split = smj_b200.dist.plan_splitters(torch.cat(allg).numpy().astype(np.uint32), world)
shard_in = []
for t in range(2):
    loc, fk, _ = recv[t]
    bnd = np.concatenate([[0], np.searchsorted(fk, split, "left"), [len(fk)]]).astype(np.int64)
    cnt = torch.from_numpy(np.diff(bnd))
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = torch.stack(allc).numpy()                              # [src][dst]
    off, total = smj_b200.dist.plan_exchange(counts, rank)
    cols = loc.shape[1]
    out_list = [torch.from_numpy(np.ascontiguousarray(loc[bnd[d]:bnd[d + 1]]).reshape(-1)) for d in range(world)]
    in_list = [torch.zeros(int(counts[s][rank]) * cols, dtype=torch.int32) for s in range(world)]
    # gloo has no all_to_all: every rank broadcasts its slices in turn
    for s in range(world):
        for d in range(world):
            buf = out_list[d].clone() if rank == s else torch.zeros(int(counts[s][d]) * cols, dtype=torch.int32)
            dist.broadcast(buf, s)
            if rank == d:
                in_list[s] = buf
    got = np.concatenate([x.numpy().reshape(-1, cols) for x in in_list]) if total else np.empty((0, cols), np.int32)
    assert got.shape[0] == total and all(off[s] == sum(counts[:s, rank]) for s in range(world))
    shard_in.append(port.sort(got, key[t]))                          # == merge of the source-ordered sorted runs (stable)
shard = port.join(shard_in[0], shard_in[1], key[0], key[1])
parts = [None] * world
dist.all_gather_object(parts, shard)
if rank == 0:
    full = np.concatenate(parts)
    want, _, _ = port.run(t1, t2, knobs["sel_col1"], knobs["sel_val1"], knobs["sel_col2"], knobs["sel_val2"], knobs["key1"], knobs["key2"])
    assert full.shape == want.shape and np.array_equal(full, want), (full.shape, want.shape)
    print("DIST_PLAN_OK", case, full.shape[0], [p.shape[0] for p in parts])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("case", ["unique", "dups", "empty_rank"])
@pytest.mark.parametrize("world", [2, 3])
def test_key_range_exchange_protocol_gloo(case, world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SMJ_ROOT=ROOT, SMJ_CASE=case, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + world * 7 + len(case)))
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DIST_PLAN_OK" in outs[0], outs[0]


def test_plan_splitters_and_exchange_units():
    import smj_b200
    s = np.array([5, 1, 9, 3, 7, 0xffffffff, 2, 8], np.uint32)
    sp = smj_b200.dist.plan_splitters(s, 4)
    srt = np.sort(s[s != 0xffffffff])
    assert list(sp) == [srt[len(srt) * b // 4] for b in (1, 2, 3)]
    assert list(smj_b200.dist.plan_splitters(np.array([], np.uint32), 3)) == [0xffffffff, 0xffffffff]
    assert len(smj_b200.dist.plan_splitters(s, 1)) == 0
    counts = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.int64)
    off, tot = smj_b200.dist.plan_exchange(counts, 1)
    assert list(off) == [0, 2, 7] and tot == 15


FABRIC_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["SMJ_ROOT"])
import torch, torch.distributed as dist
import smj_b200
from oracle import oracle

# The default multi-GPU path (csrc/smj_dist.cu "fabric" path) on CPU ranks: partition FIRST (rows grouped by destination,
# original order kept), counts gathered into a matrix, every rank derives row0 / rows / verdict with the product's own
# smj_plan_fabric, rows land at row0[dst] of the owner's receive buffer, the owner sorts stably and joins.  A receive
# capacity that is too small must give the same verdict on every rank, store nothing, and succeed after the re-sizing.
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(3)
n1, n2, c1, c2 = 30_000, 20_000, 3, 4
t1 = rng.integers(-20, 5000, size=(n1, c1)).astype(np.int32); t2 = rng.integers(-20, 5000, size=(n2, c2)).astype(np.int32)
sel, key = [(1, -5), (2, 0)], [2, 1]
tabs = [t1, t2]
port = oracle.Port()
S = 64
blocks, samples = [], []
for t in range(2):
    n = tabs[t].shape[0]
    blk = tabs[t][rank * n // world:(rank + 1) * n // world]
    blocks.append(blk)
    pos = np.minimum(((2 * np.arange(S) + 1) * max(len(blk), 1)) // (2 * S), max(len(blk) - 1, 0))
    smp = np.full(S, 0xffffffff, np.uint32)
    if len(blk):
        rows = blk[pos]
        ok = rows[:, sel[t][0]] > sel[t][1]
        smp[ok] = (rows[ok, key[t]].astype(np.int64) + 2**31).astype(np.uint32)
    samples.append(smp)
mine = torch.from_numpy(np.concatenate(samples).astype(np.int64))
allg = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allg, mine)
split = smj_b200.dist.plan_splitters(torch.cat(allg).numpy().astype(np.uint32), world)
attempts = 0
cap = [int(os.environ["SMJ_CAP"])] * 2
while True:
    attempts += 1
    recv, verdicts, needs = [], [], []
    for t in range(2):
        blk = blocks[t]
        surv = blk[blk[:, sel[t][0]] > sel[t][1]]                                   # select keeps order
        fk = (surv[:, key[t]].astype(np.int64) + 2**31).astype(np.uint32)
        bucket = np.minimum(np.searchsorted(split, fk, "right"), world - 1)         # keys >= splitter[b-1] go to b
        cnt = torch.from_numpy(np.bincount(bucket, minlength=world).astype(np.int64))
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt)
        counts = torch.stack(allc).numpy()                                           # [src][dst]
        row0, rows_mine, verdict, need = smj_b200.dist.plan_fabric(counts, rank, cap[t])
        verdicts.append(verdict); needs.append(need)
        cols = blk.shape[1]
        buf = np.full((cap[t], cols), -7, np.int32)                                  # my receive buffer
        for s in range(world):                                                       # "peer stores": src s writes its bucket d at row0_s[d]
            for d in range(world):
                part = surv[bucket == d] if rank == s else None
                k = int(counts[s][d])
                x = torch.from_numpy(np.ascontiguousarray(part).reshape(-1)) if rank == s else torch.zeros(k * cols, dtype=torch.int32)
                dist.broadcast(x, s)
                if rank == d and not verdict:
                    r0 = int(counts[:s, d].sum())
                    assert rank != s or r0 == row0[d]
                    buf[r0:r0 + k] = x.numpy().reshape(-1, cols)
        recv.append(buf[:rows_mine])
    v = torch.tensor(verdicts + needs)
    vs = [torch.zeros_like(v) for _ in range(world)]
    dist.all_gather(vs, v)
    assert all(bool((x == v).all()) for x in vs), "every rank must reach the same verdict from the same matrix"
    if not any(verdicts):
        break
    cap = [max(cap[t], needs[t] + needs[t] // 4 + 16) for t in range(2)]             # collective re-sizing, then the step again
    assert attempts < 4
shard = port.join(port.sort(recv[0], key[0]), port.sort(recv[1], key[1]), key[0], key[1])
parts = [None] * world
dist.all_gather_object(parts, shard)
if rank == 0:
    full = np.concatenate(parts)
    want, _, _ = port.run(t1, t2, sel[0][0], sel[0][1], sel[1][0], sel[1][1], key[0], key[1])
    assert full.shape == want.shape and np.array_equal(full, want), (full.shape, want.shape)
    print("FABRIC_PLAN_OK", attempts, full.shape[0], [p.shape[0] for p in parts])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,cap,attempts", [(2, 40_000, 1), (2, 100, 2), (3, 40_000, 1), (3, 3000, 2)])
def test_fabric_protocol_gloo(world, cap, attempts, tmp_path):
    script = tmp_path / "fabric_worker.py"
    script.write_text(FABRIC_WORKER)
    env = dict(os.environ, SMJ_ROOT=ROOT, SMJ_CAP=str(cap), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29800 + world * 11 + cap % 7))
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert f"FABRIC_PLAN_OK {attempts}" in outs[0], outs[0]


def test_plan_fabric_units():
    import smj_b200
    counts = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.int64)
    row0, rows, verdict, need = smj_b200.dist.plan_fabric(counts, 1, 100)
    assert list(row0) == [1, 2, 3] and rows == 15 and verdict == 0 and need == 18
    row0, rows, verdict, need = smj_b200.dist.plan_fabric(counts, 2, 17)      # rank 2's share (18) does not fit
    assert list(row0) == [5, 7, 9] and rows == 0 and verdict == 1 and need == 18
    row0, rows, verdict, need = smj_b200.dist.plan_fabric(counts, 0, 18)
    assert list(row0) == [0, 0, 0] and rows == 12 and verdict == 0
    with pytest.raises(smj_b200.SmjError):
        smj_b200.dist.plan_fabric(np.array([[-1]], np.int64), 0, 5)
