#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r3b_pytest.log; tail -5 gpurun_out/r3b_pytest.log
