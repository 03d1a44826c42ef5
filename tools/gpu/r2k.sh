for w in single-hog single-whi; do
  timeout 600 python bench.py --workload $w --steps 3 > gpurun_out/r2k_$w.json 2> gpurun_out/r2k_$w.err; tail -2 gpurun_out/r2k_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r2k_$w.json')); print('$w', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['cpu_baseline'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2k.csv python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2k_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wvm_group_kernel|wvm_deep_group" -s 27 -c 9 -o gpurun_out/grp_r2k python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2k_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aggdet_" -s 6 -c 3 -o gpurun_out/agg_r2k python bench.py --workload aggdet --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
