#!/bin/bash
# Round 2: where the C driver's CSV-parse time goes on the bundled files (SMJ_TRACE=1 prints host-side phase times)
mkdir -p gpurun_out /tmp/c1/data
T=gpurun_out/r2l
python -c "
import gzip
for n in ('g1_data1.csv','g1_data2.csv'):
    open('/tmp/c1/'+n,'wb').write(gzip.open('tests/golden/'+n+'.gz','rb').read())"
cd /tmp/c1
for i in 1 2 3; do SMJ_TRACE=1 SMJ_JSON=1 $GRAFT_REPO_ROOT/host/app g1_data1.csv g1_data2.csv 2>&1 | grep -E "trace|CSV-|^\{" ; echo; done > $GRAFT_REPO_ROOT/${T}_c1_trace.txt 2>&1
cd $GRAFT_REPO_ROOT; tail -12 ${T}_c1_trace.txt | cut -c1-400
