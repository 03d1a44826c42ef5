#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
FDB_SET_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; tail -4 gpurun_out/r3f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3f_bench.json')); print('bench', '%.4g' % d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_ms_last_step'])"
