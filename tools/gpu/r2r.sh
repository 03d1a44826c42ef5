#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_detector_set.py -x -q -m gpu 2>&1 | tail -3
run() { # name, lib
  FDB_LIB=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r2r_$1.json 2> gpurun_out/r2r_$1.err; tail -2 gpurun_out/r2r_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r2r_$1.json')); print('$1', '%.4g' % d['value'], d['ms_per_step'], d['stage1_ms'])"
}
run default ""
for v in tcp20 tca5 tcs0 tcb2; do run $v featuredetection_b200/csrc/variants/libfdb200_$v.so; done
