#!/bin/bash
# ncu --set full of one BPTT launch of the c2-shaped run (after the same command exited 0 without ncu); details + source page as text
python profiles/prof_step.py 3 bf16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dynca_bwd_tc2 -s 3 -c 1 -f -o gpurun_out/prof_bwd2 python profiles/prof_step.py 3 bf16 > gpurun_out/ncu.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu.log
ncu -i gpurun_out/prof_bwd2.ncu-rep --page details > gpurun_out/details_bwd2.txt 2>&1
ncu -i gpurun_out/prof_bwd2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_bwd2.csv 2>&1
