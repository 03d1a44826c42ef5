#!/bin/bash
mkdir -p gpurun_out
FDB_SET_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; tail -6 gpurun_out/r3e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3e_bench.json')); print('bench', '%.4g' % d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
