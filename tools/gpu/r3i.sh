#!/bin/bash
mkdir -p gpurun_out
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r3i_n8.json 2> gpurun_out/r3i_n8.err; tail -3 gpurun_out/r3i_n8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3i_n8.json')); print('n8', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['step_ms_per_rank_min_median_max'], d['e2e_step_ms_per_rank_min_median_max'], d['host_cores_per_rank'], d['facefrontal']['value'], d['facefrontal']['e2e_value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --frames-total 4096 --no-facefrontal > gpurun_out/bench_r3i_n8_strong.json 2> gpurun_out/r3i_n8_strong.err; tail -3 gpurun_out/r3i_n8_strong.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3i_n8_strong.json')); print('n8 strong', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['scaling'], d['config']['global_frames'])"
