#!/bin/bash
mkdir -p gpurun_out
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r3h_n2.json 2> gpurun_out/r3h_n2.err; tail -3 gpurun_out/r3h_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3h_n2.json')); print('n2', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['step_ms_per_rank_min_median_max'], d['host_cores_per_rank'], d['facefrontal']['value'], d['facefrontal']['e2e_value'])"
