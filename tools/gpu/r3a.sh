#!/bin/bash
# final-build evidence: full GPU test suite, default bench, launch list, ncu --set full of the window + deep kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; tail -2 gpurun_out/r3a_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3a_bench.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage1_ms'], d['cpu_baseline']['value'], d['facefrontal']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r3a.csv python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r3a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wvm_group|wvm_deep_group" -s 27 -c 9 -o gpurun_out/grp_r3a python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r3a_ncu2.log 2>&1
ls -la gpurun_out/*r3a*
