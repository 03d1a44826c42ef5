#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
FDB_WINDOW_KERNEL=tc timeout 600 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -2 gpurun_out/r2y_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['stage1_ms'])"
FDB_WINDOW_KERNEL=tc timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r2y_bench_tc.json 2> gpurun_out/r2y_bench_tc.err; tail -2 gpurun_out/r2y_bench_tc.err
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench_tc.json')); print('bench tc-only', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['stage1_ms'])"
