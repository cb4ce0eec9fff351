#!/usr/bin/env python
"""BASELINE config 1 (C1), timed on both sides on the same box: the reference's bundled data1.csv x data2.csv with
the default user.h knobs -- the one configuration the reference's own CPU program can run as it is.

    python tools/bench_c1.py [--reps 3] [--no-O0] [--csv-rows 2000000]

  ours       host/app data1.csv data2.csv (the C driver over libsmj.so; SMJ_JSON=1 gives CSV parse / H2D / device /
             D2H / CSV emit separately), median of --reps runs, end to end through result.csv
  reference  oracle/_ref/cpu_app      = `gcc -o cpu_app cpu_app.c`, the reference Makefile's recipe (Makefile:22-23)
             oracle/_ref/cpu_app_O2   = the same source at -O2
             (their own Timer brackets load_csv + select + sort + join; save_to_csv is commented out in cpu_app.c:346)
             oracle/_ref/ref_oracle   = the same stage functions with per-stage times and the CSV writer
  csv        parse / emit throughput (MB/s of text) of smj_csv_parse / smj_csv_format against the reference's load_csv /
             save_to_csv (ref_oracle's load_ms / save_ms) on the bundled files and on a larger synthetic file

The bundled files are tests/golden/g1_data{1,2}.csv.gz (byte copies of sort-merge-join/data/data{1,2}.csv; sha256 in
tests/golden/golden.json).  Prints ONE JSON line.  This is a measurement script, not the product: it may run oracle/_ref."""
import argparse
import gzip
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gunzip(name, d):
    out = os.path.join(d, name)
    with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".gz"), "rb") as g, open(out, "wb") as f:
        f.write(g.read())
    return out


def run_ours(f1, f2, d, reps, env_extra=None):
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    env = dict(os.environ, SMJ_JSON="1", **(env_extra or {}))
    rows = []
    for _ in range(reps + 1):          # the first run pays CUDA context creation and the page-in of libsmj.so
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "host", "app"), f1, f2], cwd=d, env=env, capture_output=True, text=True, timeout=600)
        wall = (time.perf_counter() - t0) * 1e3
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-2000:])
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        j["process_wall_ms"] = wall
        rows.append(j)
    rows = rows[1:]
    med = {k: statistics.median(x[k] for x in rows) for k in rows[0] if isinstance(rows[0][k], (int, float))}
    med["rows"], med["selected"] = rows[0]["rows"], rows[0]["selected"]
    # what the driver itself brackets: CSV parse + smj_run (H2D, device, D2H) + CSV emit
    med["pipeline_ms"] = med["parse_ms"] + med["h2d_ms"] + med["device_ms"] + med["d2h_ms"] + med["emit_ms"]
    return med


def run_cpu_app(binary, f1, f2):
    t0 = time.perf_counter()
    r = subprocess.run([binary, f1, f2], capture_output=True, text=True, timeout=3600)
    wall = (time.perf_counter() - t0) * 1e3
    m = re.search(r"([0-9]+\.[0-9]+)", r.stdout.split("EXEC TIME")[-1])
    return {"timer_ms": float(m.group(1)) if m else None, "process_wall_ms": wall, "rc": r.returncode}


def run_ref_oracle(binary, f1, f2, out):
    r = subprocess.run([binary, f1, f2, out], capture_output=True, text=True, timeout=3600)
    return json.loads(r.stdout.splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-O0", action="store_true", help="skip the reference Makefile's unoptimised build (~90 s)")
    ap.add_argument("--csv-rows", type=int, default=2_000_000)
    args = ap.parse_args()
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    with tempfile.TemporaryDirectory() as d:
        f1, f2 = gunzip("g1_data1.csv", d), gunzip("g1_data2.csv", d)
        text_bytes = os.path.getsize(f1) + os.path.getsize(f2)
        ours = run_ours(f1, f2, d, args.reps)
        ours_hostcsv = run_ours(f1, f2, d, args.reps, {"SMJ_CSV": "host"})
        out_bytes = os.path.getsize(os.path.join(d, "data", "result.csv"))
        ref = {}
        if os.path.exists(os.path.join(ref_dir, "cpu_app_O2")):
            ref["cpu_app_O2"] = run_cpu_app(os.path.join(ref_dir, "cpu_app_O2"), f1, f2)
            ref["stages_O2"] = run_ref_oracle(os.path.join(ref_dir, "ref_oracle"), f1, f2, os.path.join(d, "ref_result.csv"))
            if not args.no_O0:
                ref["cpu_app_O0"] = run_cpu_app(os.path.join(ref_dir, "cpu_app"), f1, f2)
        nrows = sum(ours["rows"])
        line = {
            "metric": "select+sort+merge-join throughput", "unit": "Mrows/s", "n_gpus": 1, "higher_is_better": True, "dtype": "int32",
            "data": "the reference's bundled sort-merge-join/data/data1.csv x data2.csv (tests/golden/g1_*)",
            "config": {"workload": "C1: bundled data1.csv x data2.csv, default user.h knobs, CSV in -> result.csv out", "name": "c1",
                       "rows": ours["rows"], "selected": ours["selected"], "joined": ours["joined"]},
            # same bracket as cpu_app.c's Timer (load_csv .. join, cpu_app.c:322-343) plus our CSV emit: end to end through the C driver
            "value": nrows / ours["pipeline_ms"] / 1e3, "ms_per_run": ours["pipeline_ms"],
            "ours": ours, "ours_host_csv": ours_hostcsv,
            "reference": ref,
            "vs_reference": {
                "same_config": True,
                "cpu_app_O2_timer_ms": ref.get("cpu_app_O2", {}).get("timer_ms"),
                "cpu_app_O0_timer_ms": ref.get("cpu_app_O0", {}).get("timer_ms"),
                "ratio_vs_O2": (ref["cpu_app_O2"]["timer_ms"] / ours["pipeline_ms"]) if ref.get("cpu_app_O2", {}).get("timer_ms") else None,
                "ratio_vs_O0": (ref["cpu_app_O0"]["timer_ms"] / ours["pipeline_ms"]) if ref.get("cpu_app_O0", {}).get("timer_ms") else None,
                "note": "cpu_app's Timer covers load_csv + select + sort + join (it never writes result.csv: cpu_app.c:346 is commented out); "
                        "ours covers CSV parse + H2D + device + D2H + CSV emit of result.csv",
            },
            "csv": {
                "bundled_text_MB": text_bytes / 1e6, "result_text_MB": out_bytes / 1e6,
                "gpu_parse_MBps": text_bytes / 1e6 / (ours["parse_ms"] * 1e-3), "gpu_emit_MBps": out_bytes / 1e6 / (ours["emit_ms"] * 1e-3),
                "host_parse_MBps": text_bytes / 1e6 / (ours_hostcsv["parse_ms"] * 1e-3), "host_emit_MBps": out_bytes / 1e6 / (ours_hostcsv["emit_ms"] * 1e-3),
                "reference_load_csv_MBps": (text_bytes / 1e6 / (ref["stages_O2"]["load_ms"] * 1e-3)) if "stages_O2" in ref else None,
                "reference_save_to_csv_MBps": (out_bytes / 1e6 / (ref["stages_O2"]["save_ms"] * 1e-3)) if "stages_O2" in ref and ref["stages_O2"]["save_ms"] > 0 else None,
            },
        }
        # a larger file: CSV parse / emit on the GPU are launch-latency bound at 2.7 MB
        if args.csv_rows > 0:
            import smj_b200
            from oracle import oracle
            t = smj_b200.datagen.table(args.csv_rows, 4, 5)
            big = os.path.join(d, "big.csv")
            t0 = time.perf_counter(); text = smj_b200.csv_format(t); fmt_ms = (time.perf_counter() - t0) * 1e3
            for _ in range(2):
                t0 = time.perf_counter(); text = smj_b200.csv_format(t); fmt_ms = min(fmt_ms, (time.perf_counter() - t0) * 1e3)
            open(big, "wb").write(text)
            prs_ms = 1e30
            for _ in range(3):
                t0 = time.perf_counter(); back = smj_b200.csv_parse(text); prs_ms = min(prs_ms, (time.perf_counter() - t0) * 1e3)
            assert (back == t).all()
            line["csv"]["big"] = {"rows": args.csv_rows, "cols": 4, "text_MB": len(text) / 1e6,
                                  "gpu_parse_MBps": len(text) / 1e6 / (prs_ms * 1e-3), "gpu_emit_MBps": len(text) / 1e6 / (fmt_ms * 1e-3),
                                  "note": "smj_csv_parse / smj_csv_format through ctypes, text in host memory, table to / from host (H2D and D2H inside)"}
            if oracle.have_ref():
                ref_o = oracle.Ref()
                t0 = time.perf_counter(); rt = ref_o.load_csv(big); ld = (time.perf_counter() - t0) * 1e3
                t0 = time.perf_counter(); ref_o.save_csv(os.path.join(d, "big_out.csv"), rt); sv = (time.perf_counter() - t0) * 1e3
                line["csv"]["big"].update({"reference_load_csv_MBps": len(text) / 1e6 / (ld * 1e-3), "reference_save_to_csv_MBps": len(text) / 1e6 / (sv * 1e-3)})
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
