#!/usr/bin/env python
"""bench.py -- select+sort+merge-join throughput (Mrows/s) of the B200 engine, with its HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c4|c5]

One "step" = one pass of the hot path (smj_run: select -> sort -> [exchange -> merge] -> join) over one batch of
synthetic input.  At N=1 the workload is BASELINE.json configs[1]: 10M x 10M rows, 4 int32 columns, unique int32
keys, 50 % select selectivity.  `value` is whole-job Mrows/s with inputs resident in HBM (CUDA events on the
library stream); `e2e` is the same metric through the C-ABI with HOST buffers, H2D/D2H inside the timed region.
`--impl reference` times the reference's own cpu_app.c (oracle/_ref) on the host cores on a bounded sample.
Prints ONE JSON line (rank 0)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: rows1, rows2, cols, selectivity, key_domain (0 = 3n as data/generate_data.py:9)
    "c2": dict(n1=10_000_000, n2=10_000_000, cols=4, sel=0.5, desc="synthetic 10M x 10M rows, 4 int32 cols, uniform unique int32 keys, select selectivity 50%"),
    "c3": dict(n1=200_000_000, n2=200_000_000, cols=4, sel=1.0, kind=2, key_domain=1 << 20,
               desc="synthetic 200M x 200M rows, 4 int32 cols, Zipf(1.1) keys over a 2^20-key domain (key 1 holds ~12 % of the rows: "
                    "one ~25 M-row run per side; smj_synth_table kind 2, inverse CDF), every row selected, zip semantics"),
    "c3u": dict(n1=200_000_000, n2=200_000_000, cols=4, sel=1.0, kind=1, key_domain=20_000_000,
                desc="synthetic 200M x 200M rows, 4 int32 cols, heavy duplicates (uniform over 20M keys, ~10 rows per key per side), zip semantics"),
    "c4": dict(n1=500_000_000, n2=100_000_000, cols=8, sel=0.1, desc="synthetic 500M x 100M rows, 8 int32 cols, 10% select selectivity"),
    "c5": dict(n1=2_000_000_000, n2=2_000_000_000, cols=5, sel=1.0, desc="synthetic 2B x 2B rows, 5 int32 cols, key-range partitioned"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu=0):
        self.lines, self.proc, self.gpu = [], None, gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = the upper half of the samples (idle samples between steps pull the median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def result_checksum(a):
    """Order-independent checksum of a result table: sum over rows of a hash of the row's cells, mod 2^64 (numpy)."""
    import numpy as np
    if a.size == 0:
        return 0
    h = np.zeros(a.shape[0], np.uint64)
    with np.errstate(over="ignore"):
        for c in range(a.shape[1]):
            h = (h ^ a[:, c].astype(np.int64).astype(np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
            h ^= h >> np.uint64(29)
        return int(h.sum(dtype=np.uint64))


def knobs_for(w):
    """select threshold giving the configured selectivity over keys uniform in [1, 3n]."""
    def thr(n):
        return int(3 * n * (1.0 - w["sel"])) if w["sel"] < 1.0 else 0   # sel 1.0: every key (>= 1) passes "> 0"
    tot = max(w["n1"], w["n2"])
    return thr(tot), thr(tot)


def _ref_worker(job):
    """One host core: the reference's own stage functions on its own slice of the workload (cpu_app.c keeps its join
    result in globals, so every core gets a process of its own)."""
    idx, rows, cols, n1, n2, v1, v2, steps, warmup = job
    import numpy as np
    import smj_b200
    from oracle import oracle
    ref = oracle.Ref()
    t1 = smj_b200.datagen.table(rows, cols, 1, row0=idx * rows, total_rows=n1).astype(np.int64)
    t2 = smj_b200.datagen.table(rows, cols, 2, row0=idx * rows, total_rows=n2).astype(np.int64)

    def step():
        a = ref.sort(ref.select(t1, 0, v1), 0)
        b = ref.sort(ref.select(t2, 0, v2), 0)
        return ref.join(a, b, 0, 0).shape[0]
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return time.perf_counter() - t0


def run_reference(args, w, name):
    """The reference's own CPU implementation (verbatim cpu_app.c via oracle/_ref, -O2) on a bounded sample, one
    independent copy per host core (the reference itself is single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import smj_b200
    from oracle import oracle
    if not oracle.have_ref():
        if os.path.exists("/root/reference/sort-merge-join/cpu_app.c"):
            oracle.build(ref=True)
    ncores = os.cpu_count() or 1
    v1, v2 = knobs_for(w)
    if oracle.have_ref():
        kind = "reference"
        # O(n^2) insertion sort (cpu_app.c:172-202): ~1e-9 * m^2 s per table at -O2; 65,536 rows/table keeps a step near 2-3 s.
        rows = min(w["n1"], w["n2"], 65_536)
        workers = max(1, min(ncores, w["n1"] // rows, w["n2"] // rows))
        ref_warmup = args.warmup           # the warm-up the driver asked for, as echoed in the line
        jobs = [(i, rows, w["cols"], w["n1"], w["n2"], v1, v2, args.steps, ref_warmup) for i in range(workers)]
        with mp.get_context("fork").Pool(workers) as pool:
            times = pool.map(_ref_worker, jobs)
        dt = max(times) / args.steps
        val = workers * 2 * rows / dt / 1e6
        cores = workers
        sample = (f"{workers} host cores, each running the verbatim cpu_app.c select_in_cpu + insertion_sort_in_cpu + join_in_cpu "
                  f"(gcc -O2; the reference Makefile uses no -O; the program itself is single-threaded) on its own {rows}-row slice of "
                  f"each table of the {name} workload; the sort is O(n^2), so Mrows/s falls with slice size")
    else:
        import numpy as np
        port, kind, cores = oracle.Port(), "port", 1
        ref_warmup = args.warmup
        rows = min(w["n1"], w["n2"], 4_000_000)
        t1 = smj_b200.datagen.table(rows, w["cols"], 1, total_rows=w["n1"])
        t2 = smj_b200.datagen.table(rows, w["cols"], 2, total_rows=w["n2"])
        for _ in range(ref_warmup):
            port.run(t1, t2, 0, v1, 0, v2, 0, 0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            port.run(t1, t2, 0, v1, 0, v2, 0, 0)
        dt = (time.perf_counter() - t0) / args.steps
        val = 2 * rows / dt / 1e6
        sample = f"first {rows} rows of each table; restated cpu_app (O(n log n) stable sort), single thread (oracle/_ref not built)"
    # the restated O(n log n) port on a larger sample, for context
    port = oracle.Port()
    prow = min(w["n1"], w["n2"], 2_000_000)
    p1 = smj_b200.datagen.table(prow, w["cols"], 1, total_rows=w["n1"])
    p2 = smj_b200.datagen.table(prow, w["cols"], 2, total_rows=w["n2"])
    tp = time.perf_counter()
    port.run(p1, p2, 0, v1, 0, v2, 0, 0)
    pdt = time.perf_counter() - tp
    line = {
        "impl": "reference", "metric": "select+sort+merge-join throughput", "value": val, "unit": "Mrows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": ref_warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": name, "join_mode": "zip (cpu_app.c semantics)"},
        "cpu_baseline": {"value": val, "unit": "Mrows/s", "cores": cores, "host_cores": ncores, "kind": kind, "sample": sample},
        "restated_port": {"value": 2 * prow / pdt / 1e6, "unit": "Mrows/s", "cores": 1,
                          "sample": f"first {prow} rows of each table; O(n log n) restatement of cpu_app.c (oracle/smj_oracle.c)"},
        "e2e": {"value": val, "unit": "Mrows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N>1 only: per-GPU rows fixed (weak) or total fixed")
    args = ap.parse_args()
    name = args.workload or "c2"
    w = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference(args, w, name)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        # stdout carries the one JSON line: NCCL's own debug output (the "NCCL version" banner of NCCL_DEBUG=VERSION,
        # the topology lines of NCCL_DEBUG=INFO) goes to stderr unless the caller chose a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        from bench_multi import run_multi   # one process per GPU over NCCL
        return run_multi(args, w, name)

    import numpy as np
    import smj_b200
    from smj_b200 import smj as S
    L = smj_b200.lib()
    if L.smj_device_count() < 1:
        print(json.dumps({"error": "no CUDA device; libsmj has no CPU fallback"}))
        return 1
    args.warmup = max(args.warmup, 3)
    v1, v2 = knobs_for(w)
    cfg = S.default_config(select_val1=v1, select_val2=v2)
    # both tables draw their keys from the same domain [1, 3 * max(n1, n2)] so that the join has matches at any shape
    tot = max(w["n1"], w["n2"])
    d1 = smj_b200.synth_device_table(w["n1"], w["cols"], 1, kind=w.get("kind", 0), key_domain=w.get("key_domain", 0), total_rows=tot)
    d2 = smj_b200.synth_device_table(w["n2"], w["cols"], 2, kind=w.get("kind", 0), key_domain=w.get("key_domain", 0), total_rows=tot)
    nrows = w["n1"] + w["n2"]

    def step_device():
        out, st = smj_b200.run(d1, d2, cfg=cfg, on_device=True, keep_output=True)
        L.smj_table_free(C.byref(out))
        return st

    out0, _ = smj_b200.run(d1, d2, cfg=cfg, on_device=True, keep_output=True)
    # (the multi-GPU arm prints the same pair: rows_joined + result_checksum; skipped above 2 GB of result)
    checksum = result_checksum(S.to_numpy(out0)) if out0.rows * out0.cols * 4 <= 2e9 else None
    L.smj_table_free(C.byref(out0))
    for _ in range(args.warmup):
        st = step_device()
    clocks = ClockSampler()
    clocks.start()
    L.smj_device_sync()
    t0 = time.perf_counter()
    dev_ms, launches, pass_ms, passes = 0.0, 0, 0.0, 0
    stages = {k: 0.0 for k in ("select_ms", "sort_ms", "join_ms")}
    for _ in range(args.steps):
        st = step_device()
        dev_ms += st["total_device_ms"]
        launches += st["kernel_launches"]
        pass_ms += st["sort_pass_ms_avg"] * st["sort_passes"]
        passes += st["sort_passes"]
        for k in stages:
            stages[k] += st[k]
    L.smj_device_sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk = clocks.stop()
    ms = dev_ms / args.steps
    value = nrows / ms / 1e3
    replayed = bool(st.get("graph_replayed", 0))

    # the same step on tables the library has not seen in the previous call: alternate between two table pairs.  The
    # tables' addresses reach the kernels through device cells, so the pipeline graph captured for one pair replays for
    # the other (fresh_graph_replayed); SMJ_NO_GRAPH=1 in a child process gives the kernel-by-kernel (programmatic
    # dependent launch) time, what the first call on a new SHAPE costs.
    fresh_ms, fresh_replayed, eager_ms = None, None, None
    if name == "c2" and not os.environ.get("SMJ_BENCH_NO_EAGER"):
        e1 = smj_b200.synth_device_table(w["n1"], w["cols"], 3, kind=w.get("kind", 0), key_domain=w.get("key_domain", 0), total_rows=tot)
        e2 = smj_b200.synth_device_table(w["n2"], w["cols"], 4, kind=w.get("kind", 0), key_domain=w.get("key_domain", 0), total_rows=tot)
        pairs, acc, nrep = [(d1, d2), (e1, e2)], 0.0, 0
        for i in range(4 + 2 * args.steps):
            a, b = pairs[i & 1]
            out, ste = smj_b200.run(a, b, cfg=cfg, on_device=True, keep_output=True)
            L.smj_table_free(C.byref(out))
            if i >= 4:
                acc += ste["total_device_ms"]; nrep += ste.get("graph_replayed", 0)
        fresh_ms = acc / (2 * args.steps)
        fresh_replayed = nrep == 2 * args.steps
        smj_b200.free(e1); smj_b200.free(e2)
        for _ in range(2):
            step_device()          # back to the repeated pair (the e2e leg below stages its own copies)
        if not os.environ.get("SMJ_NO_GRAPH"):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup", str(args.warmup),
                                    "--no-e2e", "--no-cpu-baseline"], env=dict(os.environ, SMJ_NO_GRAPH="1", SMJ_BENCH_NO_EAGER="1"),
                                   capture_output=True, text=True, timeout=300)
                eager_ms = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])["ms_per_step"]
            except Exception:
                eager_ms = None

    # roofline of the dominant kernel: one radix scatter pass reads 8 B and writes 8 B per selected row
    peak, peak_src = peaks()
    m_avg = (st["rows_selected"][0] + st["rows_selected"][1]) / 2.0
    m_launch = st["rows_selected"][0] + st["rows_selected"][1]     # one pass launch sorts a digit of BOTH tables' pairs
    pass_avg_ms = pass_ms / max(passes, 1)
    # bytes of an average executed pass launch: the sort runs ceil(bits(key range) / 8) passes per table (device sort plan)
    pass_bytes = st.get("sort_pass_bytes_avg", 0.0) or 16.0 * m_launch
    achieved = pass_bytes / (pass_avg_ms * 1e-3) / 1e9 if pass_avg_ms > 0 else 0.0
    dram = None   # DRAM bytes of one whole step, summed over its kernels from the committed ncu pass of this command
    dp = os.path.join(ROOT, "profiles", f"r02_{name}_dram_bytes.json")
    if os.path.exists(dp):
        try:
            dram = json.load(open(dp))
        except Exception:
            dram = None
    # DRAM bytes of one executed radix pass launch, from the same committed ncu pass (the kernel's bytes per step / the
    # launches that had work; the pair arrays are L2-resident at C2, so this is BELOW the algorithmic bytes)
    traffic = None
    if dram and "radix_pass_kernel" in dram.get("kernels", {}) and passes:
        kd = dram["kernels"]["radix_pass_kernel"]
        traffic = (kd["dram_read_bytes"] + kd["dram_write_bytes"]) / max(passes / max(args.steps, 1), 1)
    roofline = {"bound": "hbm", "kernel": f"radix_pass_kernel (onesweep scatter pass; {st['sort_passes']} launches with work per step, each over both tables' pairs)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": pass_bytes, "avg_launch_ms": pass_avg_ms,
                "launches_per_step": passes / max(args.steps, 1), "share_of_step": pass_ms / max(dev_ms, 1e-9),
                "pairs_sorted_per_launch": pass_bytes / 16.0,
                # B = the fixed algorithmic-bytes formula of the whole pipeline (four passes over every
                # selected row); "planned" = the same formula for the passes and pairs the device plan actually sorted
                "pipeline_model_bytes": st["bytes_model"], "pipeline_model_gbs": st["bytes_model"] / (ms * 1e-3) / 1e9,
                "pipeline_frac_of_peak": st["bytes_model"] / (ms * 1e-3) / 1e9 / peak,
                "pipeline_planned_bytes": st.get("bytes_planned", 0.0),
                "pipeline_planned_frac_of_peak": st.get("bytes_planned", 0.0) / (ms * 1e-3) / 1e9 / peak,
                # what the step really moved: ncu dram__bytes_read + dram__bytes_write summed over one step's kernels
                # (profiles/r02_<workload>_dram_bytes.json, taken with tools/gpu_profile.sh on this command) / this run's time
                "pipeline_dram_bytes": dram.get("dram_bytes_per_step") if dram else None,
                "pipeline_dram_frac_of_peak": (dram["dram_bytes_per_step"] / (ms * 1e-3) / 1e9 / peak) if dram else None,
                "pipeline_dram_source": dram.get("source") if dram else None,
                "note": (f"pair arrays of this workload ({8e-6 * m_avg:.0f} MB each) " +
                         ("fit the 126 MB L2: the pass is not HBM-bound here (ncu: DRAM 11 %), the fraction is of the HBM roofline the model names"
                          if 16 * m_avg < 100e6 else "exceed the 126 MB L2: the pass streams from and to HBM"))}

    # end to end through the C-ABI with pinned HOST buffers (H2D of both tables + D2H of the result inside)
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
                t0 = time.perf_counter()
            out, st2 = smj_b200.run(hp[0], hp[1], cfg=cfg, on_device=False, keep_output=True)
            d2h = out.rows * out.cols * 4
            L.smj_table_free(C.byref(out))
        e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e = {"value": nrows / e_ms / 1e3, "unit": "Mrows/s", "ms_per_step": e_ms, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "h2d_ms": st2["h2d_ms"], "d2h_ms": st2["d2h_ms"]}
        for t in hp:
            L.smj_host_free(t.data)

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle
        port = oracle.Port()
        rows = min(w["n1"], w["n2"], 10_000_000)
        t1 = smj_b200.datagen.table(rows, w["cols"], 1, total_rows=w["n1"])
        t2 = smj_b200.datagen.table(rows, w["cols"], 2, total_rows=w["n2"])
        tc = time.perf_counter()
        _, _, sms = port.run(t1, t2, 0, v1, 0, v2, 0, 0)
        cdt = time.perf_counter() - tc
        cpu = {"value": 2 * rows / cdt / 1e6, "unit": "Mrows/s", "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
               "sample": f"{rows} rows of each table ({'the full workload' if rows == w['n1'] == w['n2'] else 'a prefix'}); "
                         "restated cpu_app.c (oracle/smj_oracle.c: same select / stable sort order / zipper join, "
                         "O(n log n) sort); the verbatim O(n^2) cpu_app.c is timed by --impl reference",
               "stage_ms": {"select": sms[0], "sort": sms[1], "join": sms[2]}}

    # duplicate-key workloads: the true many-to-many COUNT of the same inputs beside the zip-mode rows_joined (smj_join_count takes
    # key-sorted tables, so the two device tables are sorted in place first -- after every timed region, they are not used again)
    many_count = None
    if w.get("kind", 0) != 0:
        S.check(L.smj_sort(C.byref(d1), 0))
        S.check(L.smj_sort(C.byref(d2), 0))
        many_count = smj_b200.join_count(d1, d2, 0, 0, mode=smj_b200.JOIN_MANY)

    line = {
        "metric": "select+sort+merge-join throughput", "value": value, "unit": "Mrows/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": name, "join_mode": "zip (cpu_app.c semantics)",
                   "rows_selected": st["rows_selected"], "rows_joined": st["rows_joined"], "many_to_many_count": many_count, "result_checksum": None if checksum is None else f"{checksum:016x}",
                   "l2": f"inputs ({w['n1'] * w['cols'] * 4 / 1e6:.0f} + {w['n2'] * w['cols'] * 4 / 1e6:.0f} MB) larger than the 126 MB L2; no explicit flush"},
        "stage_ms": {k: v / args.steps for k, v in stages.items()}, "wall_ms_per_step": wall_ms,
        "graph_replayed": replayed, "fresh_tables_ms_per_step": fresh_ms, "fresh_tables_graph_replayed": fresh_replayed,
        "eager_ms_per_step": eager_ms,
        "timing_note": "value = repeated call on the same device tables; fresh_tables_ms_per_step = alternating between two table "
                       "pairs (the pipeline graph replays for any tables of the captured shape: their addresses are read from device "
                       "cells); eager_ms_per_step = the same run with SMJ_NO_GRAPH=1, every call enqueued kernel by kernel",
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk,
    }
    print(json.dumps(line))
    smj_b200.free(d1)
    smj_b200.free(d2)
    return 0


if __name__ == "__main__":
    sys.exit(main())
