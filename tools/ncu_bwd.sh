#!/bin/bash
python profiles/prof_step.py 3 bf16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dynca_bwd_tc2 -s 3 -c 1 -f -o gpurun_out/prof_bwd2 python profiles/prof_step.py 3 bf16 > gpurun_out/ncu.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu.log
