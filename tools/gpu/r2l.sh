timeout 300 python -m pytest tests/test_gpu_aggdet.py tests/test_gpu_features.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload aggdet --steps 5 > gpurun_out/r2l_aggdet.json 2> gpurun_out/r2l_aggdet.err; tail -2 gpurun_out/r2l_aggdet.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_aggdet.json')); print('aggdet', '%.4g' % d['value'], d['frames_per_s'], '%.4g' % d['e2e']['value'], d['kernel_ms'], d['roofline']['frac'], d['cpu_baseline'])"
for w in single-hog single-whi; do
  timeout 600 python bench.py --workload $w --steps 3 > gpurun_out/r2l_$w.json 2> gpurun_out/r2l_$w.err; tail -2 gpurun_out/r2l_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r2l_$w.json')); print('$w', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['cpu_baseline'])"
done
