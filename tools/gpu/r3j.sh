#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -v 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed" | sed 's/ *\[ *[0-9]*%\]//' > gpurun_out/gpu_tests_r3j.txt; tail -2 gpurun_out/gpu_tests_r3j.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resize_kernel|pyrdown_kernel" -s 10 -c 4 -o gpurun_out/pyr_r3j python bench.py --workload facefrontal --frames 256 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3j_ncu.log 2>&1
ls -la gpurun_out/pyr_r3j.ncu-rep
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
