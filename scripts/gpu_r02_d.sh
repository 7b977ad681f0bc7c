#!/bin/bash
bash scripts/gpu_iter.sh d noncu
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_all_tests.log 2>&1; echo "all tests rc=$?" >> gpurun_out/d_all_tests.log; tail -5 gpurun_out/d_all_tests.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/d_bench.err
