#!/bin/bash
mkdir -p gpurun_out
prof() { # name, env
  env $2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wvm_group" -s 24 -c 16 --csv --log-file gpurun_out/r2w_$1.csv python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2w_$1.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2w_$1.csv')) if len(r)>5]
h=rows[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg={}
for r in rows[1:]:
    try: v=float(r[mv].replace(',',''))
    except: continue
    agg.setdefault(r[kn][-40:],[]).append(v)
tot=0
for k,v in sorted(agg.items()):
    print('$1',k,len(v),'%.3f ms'%(sum(v)/len(v)/1e6)); tot+=sum(v)/len(v)
print('$1 sum %.3f ms'%(tot/1e6))
PY
}
timeout 900 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
FDB_WINDOW_KERNEL=tc timeout 600 python -m pytest tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
prof auto FDB_X=1
prof tc FDB_WINDOW_KERNEL=tc
timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -2 gpurun_out/r2w_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['stage1_ms'])"
