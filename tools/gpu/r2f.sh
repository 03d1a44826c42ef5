timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("value %.4g e2e %.4g ms/step %.1f e2e ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print(d["stage1_ms"], d["gpu_launches"], d["host_ms_last_step"])
PY
