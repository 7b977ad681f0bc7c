#!/bin/bash
# ncu pass: launch list of whole frames + one full capture of every kernel of the voxelize+scatter chain
mkdir -p gpurun_out
python scripts/run_stage.py frame 4 > gpurun_out/plain_frame.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 120 --csv --log-file gpurun_out/launches_r02.csv \
    python scripts/run_stage.py frame 4 > gpurun_out/ncu_l.log 2>&1
python scripts/run_stage.py enc 3 > gpurun_out/plain_enc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vox_ -s 14 -c 7 -o gpurun_out/prof_enc_r02 -f \
    python scripts/run_stage.py enc 3 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_l.log gpurun_out/ncu_f.log
