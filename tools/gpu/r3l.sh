#!/bin/bash
# compute-sanitizer over the small set test that drives both window kernels (packs of 3, no early exit, 160x120 frames)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --launch-timeout 0 --error-exitcode 9 python -m pytest tests/test_detector_set.py -k "packs_of_three" -q -m gpu > gpurun_out/sanitizer_set_r3l.txt 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitizer_set_r3l.txt
