timeout 600 python bench.py --steps 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench.json"))
print("value %.4g e2e %.4g ms/step %.1f e2e ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print(d["stage1_ms"], d["gpu_launches"], d["detector_set"], d["clocks"])
print("cpu", d["cpu_baseline"])
print("ff", d.get("facefrontal"))
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2e_ref.json 2> gpurun_out/r2e_ref.err; tail -2 gpurun_out/r2e_ref.err
python -c "
import json; d=json.load(open('gpurun_out/r2e_ref.json')); print('ref', d['value'], d['ms_per_step'], d['cpu_baseline']['split_core_seconds'])"
