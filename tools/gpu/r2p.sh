#!/bin/bash
# first run of the tcgen05 window kernel: parity tests of the paths that use it, then the headline bench with both window kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
echo "--- bench tc"
timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal > gpurun_out/r2p_tc.json 2> gpurun_out/r2p_tc.err; tail -3 gpurun_out/r2p_tc.err
python -c "
import json; d=json.load(open('gpurun_out/r2p_tc.json')); print('tc', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['stage1_ms'], d['detector_set'])"
echo "--- bench mma"
FDB_WINDOW_KERNEL=mma timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal > gpurun_out/r2p_mma.json 2> gpurun_out/r2p_mma.err; tail -3 gpurun_out/r2p_mma.err
python -c "
import json; d=json.load(open('gpurun_out/r2p_mma.json')); print('mma', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['stage1_ms'], d['detector_set'])"
