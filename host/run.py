"""host/run.py -- runner with the shape of the reference's sort-merge-join/run.py:3-8
(make -> ./app data1.csv data2.csv -> print captured stdout), for the B200 engine.

    python host/run.py [data1.csv data2.csv]

The reference also runs ./cpu_app first; here that role is oracle/_ref/ref_oracle (the reference's own
cpu_app.c, test infrastructure) and it is NOT run by this script -- tests/ compare the two."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run_command(command, cwd=None):
    print(f"\n{'=' * 40}\nCommand: {' '.join(command)}")
    r = subprocess.run(command, text=True, capture_output=True, cwd=cwd)
    if r.returncode == 0:
        print("Output:\n" + (r.stdout.strip() or "No output"))
    else:
        print("Error output:\n" + (r.stderr.strip() or "No error output"))
    return r.returncode


if __name__ == "__main__":
    d1, d2 = (sys.argv[1:3] if len(sys.argv) >= 3 else ("./data/data1.csv", "./data/data2.csv"))
    rc = run_command(["make", "-s", "-C", os.path.join(ROOT, "pim-sort-merge-join_b200")])
    rc = rc or run_command(["make", "-s", "-C", HERE])
    rc = rc or run_command([os.path.join(HERE, "app"), d1, d2])
    print(f"\n{'=' * 40}")
    sys.exit(rc)
