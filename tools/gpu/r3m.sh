#!/bin/bash
# final verification of the committed build: GPU test suite, smoke(), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r3m.json 2> gpurun_out/r3m_bench.err; tail -2 gpurun_out/r3m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3m.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage1_ms'], d['cpu_baseline']['value'], d['facefrontal']['value'], d['clocks'])"
