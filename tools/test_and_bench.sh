#!/bin/bash
python -m pytest tests/test_dynca_bf16_gpu.py -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','fwd_value','ms_per_step','ms_per_step_fwd')}, d['roofline']['kernel_ms'], d['e2e']['value'])
PY
tail -3 gpurun_out/bench_quick.err
