#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r3c_pytest.log; tail -3 gpurun_out/r3c_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; tail -2 gpurun_out/r3c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3c_bench.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage1_ms'], d['host_ms_last_step'], d['facefrontal']['value'])"
