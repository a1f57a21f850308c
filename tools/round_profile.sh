#!/bin/bash
# round-end evidence: GPU test-suite, full bench line, launch list of bench.py, ncu --set full of the two step kernels
tag=${1:-r2}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_${tag}.txt; cat gpurun_out/pytest_gpu_${tag}.txt
python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.txt 2>&1; tail -4 gpurun_out/smoke_${tag}.txt
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { tail gpurun_out/bench_${tag}.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_bench.csv python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
bash tools/ncu_both.sh ${tag}
head -c 1500 gpurun_out/bench_${tag}.json; echo; cat gpurun_out/bench_ref_${tag}.json | head -c 600
