#!/usr/bin/env python
"""Regenerates tests/golden/ from the REFERENCE ITSELF (oracle/_ref = the reference's own
cpu_app.c compiled by oracle/Makefile).  Run in the authoring container, where
/root/reference exists:   python tests/golden/make_golden.py

Writes
  g1_data{1,2}.csv.gz   bundled sort-merge-join/data/data{1,2}.csv   (BASELINE config 1, byte-exact, CRLF)
  g2_data{1,2}.csv.gz   test/data/data_1.csv, data_1(1).csv          (10 k rows)
  kat{2,3,4}_data{1,2}.csv, *_expected.csv                            
  g2_expected.csv.gz, g2_{select,sort}{1,2}.sha256 inside golden.json (stage dumps)
  golden.json           knobs, row counts, first/last rows and SHA-256 of every expected result
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = os.environ.get("SMJ_REF_ROOT", "/root/reference")

KATS = {
    # name: (text1, text2, knobs)
    "kat2": ("col1,col2\n7000,1\n6000,2\n7000,3\n7000,4\n100,5\n9000,6\n",
             "col1,col2,col3\n7000,10,11\n9000,20,21\n7000,30,31\n6500,40,41\n9000,50,51\n", {}),
    "kat3": ("col1,col2\r\n70,1\r\n60,2\r\n71,3\r\n72,4\r\n10,5\r\n90,-6\r\n73,3\r\n",
             "col1,col2,col3\n-7,10,3\n-9,20,21\n7,30,3\n6,40,4\n9,50,5\n0,0,-6\n",
             dict(sel_col1=1, sel_val1=2, sel_col2=0, sel_val2=-5, key1=1, key2=2)),
    "kat4": ("col1,col2\n1,1\n",
             "col1,col2,col3\n7000,10,11\n9000,20,21\n7000,30,31\n6500,40,41\n9000,50,51\n", {}),
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as g:
        shutil.copyfileobj(f, g)


def main():
    oracle.build(ref=True)
    ref = oracle.Ref()
    meta = {"generator": "tests/golden/make_golden.py via oracle/_ref (reference cpu_app.c, unmodified)", "cases": {}}
    tmp = tempfile.mkdtemp()

    big = {
        "g1": (f"{REF}/sort-merge-join/data/data1.csv", f"{REF}/sort-merge-join/data/data2.csv"),
        "g2": (f"{REF}/test/data/data_1.csv", f"{REF}/test/data/data_1(1).csv"),
    }
    for name, (f1, f2) in big.items():
        out = os.path.join(tmp, f"{name}.csv")
        dump = os.path.join(tmp, name) if name == "g2" else None
        j, sel, ms = ref.pipeline_csv(f1, f2, out, dump_prefix=dump)
        gz_copy(f1, os.path.join(HERE, f"{name}_data1.csv.gz"))
        gz_copy(f2, os.path.join(HERE, f"{name}_data2.csv.gz"))
        lines = open(out).read().splitlines()
        case = {"knobs": {}, "selected": list(sel), "joined": j, "sha256": sha(out),
                "first_row": lines[1] if j else None, "last_row": lines[-1] if j else None,
                "input_sha256": [sha(f1), sha(f2)], "ref_stage_ms": ms}
        if dump:
            gz_copy(out, os.path.join(HERE, f"{name}_expected.csv.gz"))
            case["stages"] = {s: sha(f"{dump}_{s}.csv") for s in ("select1", "select2", "sort1", "sort2")}
        meta["cases"][name] = case
        print(name, case["selected"], j, case["sha256"])

    for name, (t1, t2, knobs) in KATS.items():
        p1, p2 = os.path.join(HERE, f"{name}_data1.csv"), os.path.join(HERE, f"{name}_data2.csv")
        with open(p1, "w", newline="") as f:
            f.write(t1)
        with open(p2, "w", newline="") as f:
            f.write(t2)
        out = os.path.join(HERE, f"{name}_expected.csv")
        j, sel, _ = ref.pipeline_csv(p1, p2, out, **knobs)
        meta["cases"][name] = {"knobs": knobs, "selected": list(sel), "joined": j, "sha256": sha(out)}
        print(name, sel, j, sha(out))

    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
