#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_set.py -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r3d.json 2> gpurun_out/r3d_bench.err; tail -2 gpurun_out/r3d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3d.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage1_ms'], d['host_ms_last_step'], d['cpu_baseline']['value'], d['facefrontal']['value'], d['facefrontal']['e2e_value'], d['gpu_launches'])"
