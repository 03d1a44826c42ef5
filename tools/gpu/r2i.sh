FDB_SET_TRACE=1 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-facefrontal > gpurun_out/r2i_a.json 2> gpurun_out/r2i_a.err; grep fdb_detector_set gpurun_out/r2i_a.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/r2i_a.json')); print('batch', d['value'], d['ms_per_step'], d['stage1_ms'])"
FDB_DEEP_BATCH=0 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-facefrontal > gpurun_out/r2i_b.json 2> gpurun_out/r2i_b.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_b.json')); print('nobatch', d['value'], d['ms_per_step'], d['stage1_ms'])"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
