#!/bin/bash
# ncu --set full (with source counters) of one forward (writing the operand history) and one BPTT launch at the c2 shape
tag=${1:-r2}
python profiles/prof_step.py 3 bf16 > gpurun_out/plain.log 2>&1 || { tail gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:dynca_.wd_tc2 -s 2 -c 2 -f -o gpurun_out/prof_${tag} python profiles/prof_step.py 3 bf16 > gpurun_out/ncu.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${tag}.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page details > gpurun_out/details_${tag}.txt 2>/dev/null
python tools/srclines.py gpurun_out/src_${tag}.csv "" 45 > gpurun_out/srclines_${tag}.txt
tail -3 gpurun_out/ncu.log; ls -la gpurun_out/ | tail -8
