#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --workload facefrontal --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3g_ff.json 2> gpurun_out/r3g_ff.err; tail -2 gpurun_out/r3g_ff.err
python -c "
import json; d=json.load(open('gpurun_out/r3g_ff.json')); print('facefrontal', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['stage1_ms'])"
