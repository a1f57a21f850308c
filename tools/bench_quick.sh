#!/bin/bash
# quick bench (no CPU baseline) + launch list
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 2500 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
