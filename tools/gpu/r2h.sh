bash tools/gpu/r2f.sh
timeout 600 python -m pytest tests/test_gpu_aggdet.py tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py --workload aggdet --steps 5 --no-cpu-baseline > gpurun_out/r2h_aggdet.json 2> gpurun_out/r2h_aggdet.err; tail -3 gpurun_out/r2h_aggdet.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_aggdet.json')); print('aggdet', d['value'], d['frames_per_s'], d['e2e']['value'], d['kernel_ms'], d['roofline']['frac'], d['detections_per_step'])"
