#!/bin/bash
# round-end evidence: full bench line, launch list, ncu --set full of the two step kernels
set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err || exit 1
python profiles/prof_step.py 6 bf16 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc2.csv python profiles/prof_step.py 6 bf16 > gpurun_out/ncu1.log 2>&1
python profiles/prof_step.py 3 bf16 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dynca_.wd_tc2 -s 2 -c 2 -f -o gpurun_out/prof_r1_tc2 python profiles/prof_step.py 3 bf16 > gpurun_out/ncu2.log 2>&1
tail -c 600 gpurun_out/bench_full.json
