#!/bin/bash
# ncu evidence for every kernel: (1) the first invocation of each kernel of the library at reduced sizes (full set),
# (2) every kernel of one whole frame at benchmark size (full set), (3) the launch list of whole frames.
# Reports without imported source so that they stay far below the 64 MiB that travels back.
mkdir -p gpurun_out
python scripts/run_all_kernels.py small > gpurun_out/plain_all.log 2>&1 &&
ncu --set full --clock-control none --kernel-id :::1 -o gpurun_out/prof_all_small_r02 -f python scripts/run_all_kernels.py small > gpurun_out/ncu_all.log 2>&1
tail -n 2 gpurun_out/ncu_all.log
python scripts/run_stage.py frame 2 > gpurun_out/plain_frame.log 2>&1 &&
ncu --set full --clock-control none -s 20 -c 20 -o gpurun_out/prof_frame_r02 -f python scripts/run_stage.py frame 2 > gpurun_out/ncu_f.log 2>&1
tail -n 2 gpurun_out/ncu_f.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 130 --csv --log-file gpurun_out/launches_r02.csv \
    python scripts/run_stage.py frame 6 > gpurun_out/ncu_l.log 2>&1
tail -n 2 gpurun_out/ncu_l.log
ls -la gpurun_out/
