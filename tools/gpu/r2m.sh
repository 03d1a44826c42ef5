timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in single-hog single-whi; do
  timeout 600 python bench.py --workload $w --steps 3 --no-cpu-baseline > gpurun_out/r2m_$w.json 2> gpurun_out/r2m_$w.err; tail -2 gpurun_out/r2m_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r2m_$w.json')); print('$w', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"
done
