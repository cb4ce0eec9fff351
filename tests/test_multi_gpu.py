"""The real N>1 path: torchrun world_size 2 (and 4 when available) on one box, one process per GPU, NCCL exchange
inside libsmj.so.  Every rank checks nothing itself; rank 0 gathers the shards and compares their concatenation with
the oracle's single-process result, bit for bit.  Skipped on boxes with fewer GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["SMJ_ROOT"])
import torch, torch.distributed as dist
import smj_b200
from oracle import oracle
rank, world, local = smj_b200.dist.init()
for case in os.environ["SMJ_CASE"].split(","):
    rng = np.random.default_rng(5)
    if case == "unique":
        n1, n2, c1, c2 = 300_000, 200_000, 4, 4
        t1 = smj_b200.datagen.table(n1, c1, 1); t2 = smj_b200.datagen.table(n2, c2, 2, total_rows=n1)
        kn = dict(select_col1=0, select_val1=n1, select_col2=0, select_val2=n1 // 2, join_key1=0, join_key2=0)
    elif case == "dups":
        n1, n2, c1, c2 = 150_000, 120_000, 3, 5
        t1 = rng.integers(-50, 400, size=(n1, c1)).astype(np.int32); t2 = rng.integers(-50, 400, size=(n2, c2)).astype(np.int32)
        kn = dict(select_col1=1, select_val1=-5, select_col2=2, select_val2=0, join_key1=2, join_key2=1)
    elif case == "zipf":   # one key holds ~10 % of the rows: its whole run must land on one rank
        n1, n2, c1, c2 = 200_000, 150_000, 4, 4
        t1 = smj_b200.datagen.zipf_table(n1, c1, 31); t2 = smj_b200.datagen.zipf_table(n2, c2, 32)
        kn = dict(select_col1=0, select_val1=5000, select_col2=0, select_val2=5000, join_key1=0, join_key2=0)
    else:   # tiny and skewed: some ranks own nothing
        n1, n2, c1, c2 = 37, 11, 2, 2
        t1 = rng.integers(0, 5, size=(n1, c1)).astype(np.int32); t2 = rng.integers(0, 5, size=(n2, c2)).astype(np.int32)
        kn = dict(select_col1=0, select_val1=-1, select_col2=0, select_val2=-1, join_key1=0, join_key2=0)
    b1 = t1[rank * n1 // world:(rank + 1) * n1 // world]
    b2 = t2[rank * n2 // world:(rank + 1) * n2 // world]
    if case == "unique":
        # a smaller problem first: the second run needs larger receive buffers on every rank, so the peer mappings of the
        # first one are stale and must be re-exchanged (CUDA-IPC path)
        s1, s2 = smj_b200.datagen.table(20_000, c1, 5), smj_b200.datagen.table(10_000, c2, 6, total_rows=20_000)
        sk = dict(select_col1=0, select_val1=100, select_col2=0, select_val2=100, join_key1=0, join_key2=0)
        shard, _ = smj_b200.run(s1[rank * 20_000 // world:(rank + 1) * 20_000 // world], s2[rank * 10_000 // world:(rank + 1) * 10_000 // world],
                                nr_gpus=world, **sk)
        parts = [None] * world
        dist.all_gather_object(parts, shard)
        if rank == 0:
            want, _, _ = oracle.Port().run(s1, s2, 0, 100, 0, 100, 0, 0)
            assert np.array_equal(np.concatenate(parts), want), "small warm-up problem"
    for on_device in (False, True):
        if on_device:
            a, b = smj_b200.device_table(b1), smj_b200.device_table(b2)
            shard, st = smj_b200.run(a, b, on_device=True, nr_gpus=world, **kn)
            smj_b200.free(a); smj_b200.free(b)
        else:
            shard, st = smj_b200.run(b1, b2, nr_gpus=world, **kn)
        parts = [None] * world
        dist.all_gather_object(parts, shard)
        if rank == 0:
            full = np.concatenate(parts)
            want, sel, _ = oracle.Port().run(t1, t2, kn["select_col1"], kn["select_val1"], kn["select_col2"], kn["select_val2"],
                                              kn["join_key1"], kn["join_key2"])
            assert full.shape == want.shape and np.array_equal(full, want), (case, full.shape, want.shape)
            print("MULTI_OK", case, on_device, [p.shape[0] for p in parts], st["bytes_nvlink"], flush=True)
smj_b200.lib().smj_shutdown()
dist.barrier()
dist.destroy_process_group()
'''


def _ngpus():
    import smj_b200
    return smj_b200.lib().smj_device_count()


# path: "peer" (default) = the fabric path: select+partition, mailboxes + exchange stores over peer memory, no host wait;
#       "peer1" = the same on one stream; "overflow" = receive buffers far too small, so every step takes the collective
#       verdict -> re-size -> re-run route; "nccl" = same partitioning, NCCL collectives + grouped ncclSend/ncclRecv;
#       "merge" = sort first, exchange, merge-path merge tree
@pytest.mark.parametrize("world,path", [(2, "peer"), (2, "peer1"), (2, "overflow"), (2, "nccl"), (2, "merge"),
                                        (4, "peer"), (4, "overflow"), (4, "nccl"), (4, "merge"), (8, "peer"), (8, "merge")])
def test_key_range_join_matches_single_process_oracle(world, path, tmp_path):
    case = "unique,dups,zipf,tiny"   # one launch per (world, path): the cases share the process group
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SMJ_ROOT=ROOT, SMJ_CASE=case)
    env.pop("SMJ_DIST_MODE", None)
    env.pop("SMJ_DIST_EXCHANGE", None)
    env.pop("SMJ_DIST_STREAMS", None)
    env.pop("SMJ_DIST_CAP_PCT", None)
    if path == "peer1":
        env["SMJ_DIST_STREAMS"] = "1"
    if path == "overflow":
        env["SMJ_DIST_CAP_PCT"] = "10"
    if path == "nccl":
        env["SMJ_DIST_EXCHANGE"] = "nccl"
    if path == "merge":
        env["SMJ_DIST_MODE"] = "merge"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + world), str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_OK") == 8, r.stdout[-2000:]


LOCAL = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["SMJ_ROOT"])
import smj_b200
from oracle import oracle
G = int(os.environ["SMJ_G"])
rng = np.random.default_rng(7)
port = oracle.Port()
cases = []
n1, n2 = 300_000, 200_000
cases.append(("unique", smj_b200.datagen.table(n1, 4, 1), smj_b200.datagen.table(n2, 4, 2, total_rows=n1),
              dict(select_col1=0, select_val1=n1, select_col2=0, select_val2=n1 // 2, join_key1=0, join_key2=0)))
cases.append(("dups", rng.integers(-50, 400, size=(150_000, 3)).astype(np.int32), rng.integers(-50, 400, size=(120_000, 5)).astype(np.int32),
              dict(select_col1=1, select_val1=-5, select_col2=2, select_val2=0, join_key1=2, join_key2=1)))
cases.append(("zipf", smj_b200.datagen.zipf_table(200_000, 4, 31), smj_b200.datagen.zipf_table(150_000, 4, 32),
              dict(select_col1=0, select_val1=5000, select_col2=0, select_val2=5000, join_key1=0, join_key2=0)))
cases.append(("wide", smj_b200.datagen.table(400_000, 8, 11), smj_b200.datagen.table(80_000, 8, 12, total_rows=400_000),      # C4's shape: 8 columns, 10 % select
              dict(select_col1=0, select_val1=1_080_000, select_col2=0, select_val2=1_080_000, join_key1=0, join_key2=0)))
cases.append(("tiny", rng.integers(0, 5, size=(37, 2)).astype(np.int32), rng.integers(0, 5, size=(11, 2)).astype(np.int32),
              dict(select_col1=0, select_val1=-1, select_col2=0, select_val2=-1, join_key1=0, join_key2=0)))
cases.append(("empty", np.zeros((0, 3), np.int32), rng.integers(0, 5, size=(11, 2)).astype(np.int32),
              dict(select_col1=0, select_val1=-1, select_col2=0, select_val2=-1, join_key1=0, join_key2=0)))
for name, t1, t2, kn in cases:
    want, sel, _ = port.run(t1, t2, kn["select_col1"], kn["select_val1"], kn["select_col2"], kn["select_val2"], kn["join_key1"], kn["join_key2"])
    for rep in range(2):     # the second call repeats the first: graph replay of the local pipelines
        got, st = smj_b200.run(t1, t2, nr_gpus=G, **kn)
        assert got.shape == want.shape and np.array_equal(got, want), (name, rep, got.shape, want.shape)
        assert st["rows_selected"] == list(sel), (name, st["rows_selected"], sel)
        assert st["rows_joined"] == want.shape[0]
    if t1.shape[0]:
        a, b = smj_b200.device_table(t1), smj_b200.device_table(t2)     # device tables on GPU 0: blocks travel to the other ranks
        got, st = smj_b200.run(a, b, nr_gpus=G, **kn)
        smj_b200.free(a); smj_b200.free(b)
        assert np.array_equal(got, want), (name, "device tables")
    print("LOCAL_OK", name, G, want.shape[0], flush=True)
# back to one GPU in the same process
got, st = smj_b200.run(cases[0][1], cases[0][2], nr_gpus=1, **cases[0][3])
want, _, _ = port.run(cases[0][1], cases[0][2], 0, n1, 0, n1 // 2, 0, 0)
assert np.array_equal(got, want)
print("LOCAL_OK single", flush=True)
smj_b200.lib().smj_shutdown()
'''


@pytest.mark.parametrize("world,variant", [(2, "default"), (2, "overflow"), (4, "default"), (8, "default"),
                                           (2, "one_gpu"), (3, "one_gpu"), (4, "one_gpu_overflow"), (8, "one_gpu")])
def test_one_process_drives_all_gpus(world, variant, tmp_path):
    """smj_run with nr_gpus = G from ONE process (what host/app does with SMJ_NR_GPUS=G): rows dealt to G devices, the
    exchange over peer memory, ONE result table in key order -- bit-identical to the oracle's single-process result.
    The one_gpu variants put all G ranks on device 0 (SMJ_RANKS_ON_ONE_GPU=1): the same mailboxes, bucket routing, receive
    buffers and verdicts, rank against rank on concurrent streams, on a box with a single GPU."""
    if _ngpus() < world and not variant.startswith("one_gpu"):
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "local.py"
    script.write_text(LOCAL)
    env = dict(os.environ, SMJ_ROOT=ROOT, SMJ_G=str(world))
    for k in ("SMJ_DIST_MODE", "SMJ_DIST_EXCHANGE", "SMJ_DIST_STREAMS", "SMJ_DIST_CAP_PCT"):
        env.pop(k, None)
    if variant.endswith("overflow"):
        env["SMJ_DIST_CAP_PCT"] = "10"
    if variant.startswith("one_gpu"):
        env["SMJ_RANKS_ON_ONE_GPU"] = "1"
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("LOCAL_OK") == 7, r.stdout[-2000:]
