run() { # name lib
  FDB_LIB=$2 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-facefrontal > gpurun_out/r2j_$1.json 2> gpurun_out/r2j_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r2j_$1.json')); print('$1', '%.4g' % d['value'], '%.1f' % d['ms_per_step'], {k: round(v,1) for k,v in d['stage1_ms'].items()}, d['host_ms_last_step']['wait_stage1'])"
}
V=featuredetection_b200/csrc/variants
run base featuredetection_b200/csrc/libfdb200.so
run ku2 $V/libfdb200_ku2.so
run pack3 $V/libfdb200_pack3.so
run pack4 $V/libfdb200_pack4.so
run run18 $V/libfdb200_run18.so
FDB_LIB=$V/libfdb200_pack4.so timeout 300 python -m pytest tests/test_detector_set.py -x -q -m gpu 2>&1 | tail -2
