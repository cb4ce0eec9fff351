"""bench_multi.py -- the N>1 arm of bench.py: one process per GPU (torchrun), key-range partitioned join.

Weak scaling: every rank holds a 10M-row block of each table of a virtual N x 10M-row pair (unique keys in
[1, 3*N*10M], 50 % select), i.e. per-GPU input is BASELINE configs[1] whatever N is; the exchange step (grouped
ncclSend/ncclRecv inside libsmj.so) moves (N-1)/N of the selected rows.  Timing is on the device (CUDA events inside
smj_run, library stream), max over ranks; rank 0 prints the JSON line."""
import ctypes as C
import json
import os
import time


def run_multi(args, w, name):
    import smj_b200
    from smj_b200 import smj as S
    from bench import ClockSampler, peaks, knobs_for
    rank, world, local = smj_b200.dist.init()
    L = smj_b200.lib()
    G = world
    n1, n2, cols = w["n1"], w["n2"], w["cols"]
    if args.scaling == "strong":
        n1, n2 = n1 // G, n2 // G
    tot1, tot2 = n1 * G, n2 * G
    wv = dict(w, n1=tot1, n2=tot2)
    v1, v2 = knobs_for(wv)
    cfg = S.default_config(select_val1=v1, select_val2=v2, nr_gpus=G)
    # both tables draw their keys from the same domain [1, 3 * max(rows)] (as the single-GPU arm does), so that the join has
    # matches at any shape; rank r holds rows [r * n, (r + 1) * n) of each virtual table
    dom_rows = max(tot1, tot2)
    kind, kdom = w.get("kind", 0), w.get("key_domain", 0)
    d1 = smj_b200.synth_device_table(n1, cols, 1, kind=kind, key_domain=kdom, row0=rank * n1, total_rows=dom_rows)
    d2 = smj_b200.synth_device_table(n2, cols, 2, kind=kind, key_domain=kdom, row0=rank * n2, total_rows=dom_rows)
    args.warmup = max(args.warmup, 3)

    def step():
        out, st = smj_b200.run(d1, d2, cfg=cfg, on_device=True, keep_output=True)
        L.smj_table_free(C.byref(out))
        return st

    # order-independent checksum of the whole result (sum over rows of a hash of the row's cells, mod 2^64) and its row
    # count, summed over the ranks: the same virtual tables must give the same pair at every N (strong scaling) -- a
    # cheap end-to-end parity check at sizes no CPU oracle reaches (bench.py prints the same pair at N=1)
    from bench import result_checksum

    def checked_step():
        out, st_ = smj_b200.run(d1, d2, cfg=cfg, on_device=True, keep_output=True)
        big_ = smj_b200.dist.max_over_ranks(out.rows * out.cols * 4) > 2e9     # (skipped where a shard would not fit comfortably in host memory)
        ck = 0 if big_ else result_checksum(smj_b200.smj.to_numpy(out))
        rows_ = int(smj_b200.dist.sum_over_ranks(out.rows))
        L.smj_table_free(C.byref(out))
        ck_lo = int(smj_b200.dist.sum_over_ranks(ck & 0xffffff))                 # (float64 all-reduce: 24-bit pieces stay exact; carries kept)
        ck_mid = int(smj_b200.dist.sum_over_ranks((ck >> 24) & 0xffffff))
        ck_hi = int(smj_b200.dist.sum_over_ranks(ck >> 48))
        return big_, (ck_lo + (ck_mid << 24) + (ck_hi << 48)) & 0xffffffffffffffff, rows_

    big, checksum, rows_first = checked_step()      # the very first step (eager launches, fresh receive buffers)

    for _ in range(args.warmup):
        st = step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    smj_b200.dist.barrier()
    t0 = time.perf_counter()
    dev_ms, launches, pass_ms, passes = 0.0, 0, 0.0, 0
    stages = {k: 0.0 for k in ("select_ms", "sort_ms", "exchange_ms", "merge_ms", "join_ms")}
    nvlink = 0.0
    for _ in range(args.steps):
        st = step()
        dev_ms += st["total_device_ms"]
        launches += st["kernel_launches"]
        pass_ms += st["sort_pass_ms_avg"] * st["sort_passes"]
        passes += st["sort_passes"]
        nvlink += st["bytes_nvlink"]
        for k in stages:
            stages[k] += st[k]
    smj_b200.dist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk = clocks.stop() if rank == 0 else None
    _, checksum_last, rows_last = checked_step()    # ... and a step after the timed ones (graph replay): must be the same result
    ms = smj_b200.dist.max_over_ranks(dev_ms / args.steps)
    wall_ms = smj_b200.dist.max_over_ranks(wall_ms)
    launches = int(smj_b200.dist.sum_over_ranks(launches))
    joined = int(smj_b200.dist.sum_over_ranks(st["rows_joined"]))
    sel = [int(smj_b200.dist.sum_over_ranks(st["rows_selected"][t])) for t in range(2)]
    nvlink_max = smj_b200.dist.max_over_ranks(nvlink / args.steps)
    stage_max = {k: smj_b200.dist.max_over_ranks(v / args.steps) for k, v in stages.items()}
    nrows = tot1 + tot2
    value = nrows / ms / 1e3

    # end to end through the C-ABI with pinned HOST buffers on every rank
    e2e = None
    if not args.no_e2e:
        hp = []
        for d in (d1, d2):
            p = C.c_void_p()
            S.check(L.smj_host_alloc(C.byref(p), d.rows * d.cols * 4))
            S.check(L.smj_memcpy_d2h(p, d.data, d.rows * d.cols * 4))
            hp.append(S.Table(p.value, d.rows, d.cols, 0))
        h2d = sum(t.rows * t.cols * 4 for t in hp)
        d2h = 0
        for i in range(2 + args.steps):
            if i == 2:
                smj_b200.dist.barrier()
                t0 = time.perf_counter()
            out, st2 = smj_b200.run(hp[0], hp[1], cfg=cfg, on_device=False, keep_output=True)
            d2h = out.rows * out.cols * 4
            L.smj_table_free(C.byref(out))
        smj_b200.dist.barrier()
        e_ms = smj_b200.dist.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        e2e = {"value": nrows / e_ms / 1e3, "unit": "Mrows/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(smj_b200.dist.sum_over_ranks(h2d)), "d2h_bytes_per_step": int(smj_b200.dist.sum_over_ranks(d2h))}
        for t in hp:
            L.smj_host_free(t.data)

    if rank == 0:
        peak, peak_src = peaks()
        m_launch = (sel[0] + sel[1]) / G                 # one pass launch sorts a digit of both tables' pairs on one GPU
        pass_avg_ms = pass_ms / max(passes, 1)
        pass_bytes = st.get("sort_pass_bytes_avg", 0.0) or 16.0 * m_launch   # rank 0's executed passes (sort plan)
        achieved = pass_bytes / (pass_avg_ms * 1e-3) / 1e9 if pass_avg_ms > 0 else 0.0
        line = {
            "metric": "select+sort+merge-join throughput", "value": value, "unit": "Mrows/s", "n_gpus": G, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": (f"{w['desc']} per GPU (weak scaling)" if args.scaling == "weak" else f"{w['desc']} in all (strong scaling)") +
                                   f": {tot1} x {tot2} rows over {G} GPUs ({n1} x {n2} per GPU), key-range partitioned, sample / count mailboxes and the "
                                   "exchange stores over NVLink peer memory (no NCCL call, no host wait in a step)", "name": name, "join_mode": "zip (cpu_app.c semantics)",
                       "rows_selected": sel, "rows_joined": joined, "result_checksum": None if big else f"{checksum:016x}",
                       "result_checksum_after_timed_steps": None if big else f"{checksum_last:016x}", "rows_joined_first_and_last_step": [rows_first, rows_last],
                       "parallelism": f"key-range x{G}",
                       "exchange": os.environ.get("SMJ_DIST_EXCHANGE", "fabric") + ("/" + os.environ["SMJ_DIST_MODE"] if os.environ.get("SMJ_DIST_MODE") else ""),
                       "l2": f"per-GPU inputs ({n1 * cols * 4 / 1e6:.0f} + {n2 * cols * 4 / 1e6:.0f} MB) larger than the 126 MB L2; no explicit flush"},
            "stage_ms": stage_max, "wall_ms_per_step": wall_ms, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": f"radix_pass_kernel (onesweep scatter pass; {st['sort_passes']} launches with work per step, each over both tables' pairs)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": pass_bytes, "avg_launch_ms": pass_avg_ms,
                         "nvlink_bytes_sent_per_gpu": nvlink_max,
                         # bytes the busiest rank stored into peer memory / the exchange window (table 2's partition pass runs inside it)
                         "nvlink_gbs_per_gpu": nvlink_max / (stage_max["exchange_ms"] * 1e-3) / 1e9 if stage_max["exchange_ms"] > 0 else None,
                         "nvlink_peak_gbs": 770.0, "nvlink_peak_source": "peer copy rate measured on this pool in round 1 (900 GB/s nominal per direction)"},
            "cpu_baseline": None, "e2e": e2e, "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    smj_b200.free(d1)
    smj_b200.free(d2)
    L.smj_shutdown()
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
    return 0
