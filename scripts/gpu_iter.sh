#!/bin/bash
# one development iteration on the GPU box: voxelizer parity, per-kernel event times, ncu launch list + full capture
# usage: gpu_iter.sh TAG [notests] [noncu]
TAG=${1:-x}
mkdir -p gpurun_out
if [[ "$*" != *notests* ]]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "voxelize or frame_pipeline or pfn or custom_voxelizer" > gpurun_out/${TAG}_vox_tests.log 2>&1
  echo "vox tests rc=$?" >> gpurun_out/${TAG}_vox_tests.log
  tail -15 gpurun_out/${TAG}_vox_tests.log
fi
for mode in "" given uniform; do
  timeout 300 python scripts/prof_events.py enc $mode > gpurun_out/${TAG}_prof_enc_${mode:-refl}.log 2>&1
  cat gpurun_out/${TAG}_prof_enc_${mode:-refl}.log
done
timeout 300 python scripts/inflight.py 24 > gpurun_out/${TAG}_inflight.log 2>&1; tail -8 gpurun_out/${TAG}_inflight.log
if [[ "$*" != *noncu* ]]; then
  python scripts/run_stage.py enc 3 > gpurun_out/plain_enc.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:vox_ -s 12 -c 6 -o gpurun_out/prof_enc_${TAG} -f \
      python scripts/run_stage.py enc 3 > gpurun_out/ncu_f.log 2>&1
  tail -n 3 gpurun_out/ncu_f.log
fi
