#!/bin/bash
# ncu --set full (with source counters) of one ConditionedNCA forward + BPTT launch at the c4 shape (B=64 to keep it short)
tag=${1:-enc}
python tools/perf_enc.py 64 x > gpurun_out/plain_enc.log 2>&1 || { tail gpurun_out/plain_enc.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:enc_bwd_tc -s 3 -c 1 -f -o gpurun_out/prof_${tag} python tools/perf_enc.py 64 x > gpurun_out/ncu_enc.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${tag}.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page details > gpurun_out/details_${tag}.txt 2>/dev/null
python tools/srclines.py gpurun_out/src_${tag}.csv "" 40 > gpurun_out/srclines_${tag}.txt
tail -2 gpurun_out/ncu_enc.log
