"""One-process-per-GPU plumbing for the key-range partitioned path (include/smj.h: smj_init_dist).

torch.distributed is used ONLY to rendezvous: rank 0 asks libsmj for an ncclUniqueId and broadcasts the 128 bytes;
every data-path collective (sample all-gather, count all-gather, the grouped send/recv all-to-all of both tables)
runs inside libsmj.so on its own NCCL communicator.  Launch with torchrun / torch.distributed.run."""
import ctypes as C
import os

from . import smj as S


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def init(cfg=None, backend=None):
    """Initialises torch.distributed (if needed) and libsmj's distributed mode.  Returns (rank, world, local_rank)."""
    import torch                      # first: libsmj's lazy dlopen("libnccl.so.2") then shares torch's NCCL
    import torch.distributed as dist
    rank, world, local = env()
    L = S.lib()
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    dev = torch.device("cuda", local) if (use_cuda and dist.get_backend() == "nccl") else torch.device("cpu")
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        S.check(L.smj_dist_unique_id(idbuf))
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    cfg = cfg or S.default_config()
    cfg.nr_gpus = world
    S.check(L.smj_init_dist(C.byref(cfg), rank, world, local, raw))
    return rank, world, local


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x):
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ---------------------------------------------------------------- host-only planning (ctypes over the C functions)
def plan_splitters(samples, world):
    import numpy as np
    s = np.ascontiguousarray(samples, dtype=np.uint32)
    out = np.zeros(max(world - 1, 1), np.uint32)
    S.check(S.lib().smj_plan_splitters(s.ctypes.data, s.size, world, out.ctypes.data))
    return out[:world - 1]


def plan_exchange(counts, me):
    import numpy as np
    c = np.ascontiguousarray(counts, dtype=np.int64)
    world = c.shape[0]
    off = np.zeros(world, np.int64)
    tot = C.c_int64()
    S.check(S.lib().smj_plan_exchange(c.ctypes.data, world, me, off.ctypes.data, C.byref(tot)))
    return off, tot.value


def plan_fabric(counts, me, cap_rows):
    """smj_plan_fabric: (row0[world], rows_mine, verdict, need_rows) from the world x world count matrix."""
    import numpy as np
    c = np.ascontiguousarray(counts, dtype=np.int64)
    world = c.shape[0]
    row0 = np.zeros(world, np.int64)
    rows, verdict, need = C.c_int64(), C.c_int(), C.c_int64()
    S.check(S.lib().smj_plan_fabric(c.ctypes.data, world, me, int(cap_rows), row0.ctypes.data, C.byref(rows), C.byref(verdict), C.byref(need)))
    return row0, rows.value, verdict.value, need.value
