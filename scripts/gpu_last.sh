#!/bin/bash
# tests + smoke + bench + the whole-frame ncu table (usage: gpu_last.sh TAG [bench args])
TAG=${1:-last}; shift
bash scripts/gpu_full.sh $TAG "$@"
python scripts/run_stage.py frame 2 > gpurun_out/plain_frame.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -s 18 -c 18 -o /tmp/prof_frame -f python scripts/run_stage.py frame 2 > gpurun_out/ncu_f.log 2>&1
tail -n 1 gpurun_out/ncu_f.log
python scripts/ncu_table.py /tmp/prof_frame.ncu-rep gpurun_out/${TAG}_ncu_frame_full > /dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 115 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/run_stage.py frame 6 > gpurun_out/ncu_l.log 2>&1
cat gpurun_out/${TAG}_ncu_frame_full.md | cut -c1-150
