#!/bin/bash
# quick GPU check: network / full-size / round-2 suites + the two training bench lines
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests/test_gpu_network.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -q --tb=short -rP > gpurun_out/${TAG}_tests.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_tests.log | tail -8
grep -E "FULLSIZE|TRAJECTORY" gpurun_out/${TAG}_tests.log | cut -c1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_2d.json 2> gpurun_out/${TAG}_bench_2d.err; tail -c 300 gpurun_out/${TAG}_bench_2d.err
python bench.py --config 1d --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1d.json 2> gpurun_out/${TAG}_bench_1d.err
for c in 2d 1d; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1])
    print('$c', round(d['value'],1), 'clips/s', round(d['ms_per_step'],3), 'ms', {k: v['ms'] for k, v in d['phases_ms'].items()})
except Exception as e: print('$c bench failed', e)
PY
done
