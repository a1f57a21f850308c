#!/bin/bash
# round-end verification on one B200 (run under gpurun): full GPU parity suite, smoke, both bench arms, callers bench + its ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref_r1.json
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
python tools/bench_callers.py > gpurun_out/callers_r1.jsonl 2> gpurun_out/callers.err; tail -3 gpurun_out/callers.err; cat gpurun_out/callers_r1.jsonl
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r1_callers.csv python tools/bench_callers.py --quick > gpurun_out/ncu_callers.log 2>&1
grep -c . gpurun_out/launches_r1_callers.csv
