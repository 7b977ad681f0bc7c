#!/bin/bash
# round 2, first GPU pass: parity of the new voxelizer + per-kernel event times
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "voxelize or frame_pipeline or pfn" > gpurun_out/a_vox_tests.log 2>&1
echo "vox tests rc=$?" >> gpurun_out/a_vox_tests.log
tail -30 gpurun_out/a_vox_tests.log
for mode in "" features unfused; do
  timeout 300 python scripts/prof_events.py enc $mode > gpurun_out/a_prof_enc_${mode:-scatter}.log 2>&1
  cat gpurun_out/a_prof_enc_${mode:-scatter}.log
done
timeout 300 python scripts/prof_events.py enc given > gpurun_out/a_prof_enc_given.log 2>&1; cat gpurun_out/a_prof_enc_given.log
timeout 300 python scripts/prof_events.py enc uniform > gpurun_out/a_prof_enc_uniform.log 2>&1; cat gpurun_out/a_prof_enc_uniform.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/a_all_tests.log
tail -40 gpurun_out/a_all_tests.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
