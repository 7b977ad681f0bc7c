#!/bin/bash
# round-end evidence in one call: all gpu tests, smoke, bench, then the ncu tables (whole frame at benchmark size and the
# first invocation of every kernel at reduced sizes; reports stay on the box, the tables and the raw csv travel back)
# and the launch list.  usage: gpu_final.sh TAG [bench args]
TAG=${1:-final}; shift
mkdir -p gpurun_out
bash scripts/gpu_full.sh $TAG "$@"
python scripts/run_stage.py frame 2 > gpurun_out/plain_frame.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -s 20 -c 20 -o /tmp/prof_frame -f python scripts/run_stage.py frame 2 > gpurun_out/ncu_f.log 2>&1
tail -n 1 gpurun_out/ncu_f.log
python scripts/ncu_table.py /tmp/prof_frame.ncu-rep gpurun_out/${TAG}_ncu_frame_full > /dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 130 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/run_stage.py frame 6 > gpurun_out/ncu_l.log 2>&1
# the clipped NMS kernels at 20k boxes: filter, level-2 mask and level-2 sweep of the first call, per pair test
for m in rot_bev box3d; do
  python scripts/run_nms_mode.py $m 2 > gpurun_out/plain_nms_$m.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none -k regex:"nms_mask_clip|nms_sweep|nms_filter_clip" -s 2 -c 3 -o /tmp/prof_nms_$m -f python scripts/run_nms_mode.py $m 2 > gpurun_out/ncu_nms_$m.log 2>&1
  python scripts/ncu_table.py /tmp/prof_nms_$m.ncu-rep gpurun_out/${TAG}_ncu_nms20k_$m > /dev/null
done
python scripts/run_all_kernels.py small > gpurun_out/plain_all.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --kernel-id :::1 -o /tmp/prof_all_small -f python scripts/run_all_kernels.py small > gpurun_out/ncu_all.log 2>&1
tail -n 1 gpurun_out/plain_all.log gpurun_out/ncu_all.log
python scripts/ncu_table.py /tmp/prof_all_small.ncu-rep gpurun_out/${TAG}_ncu_all_kernels_small > /dev/null
ncu -i /tmp/prof_all_small.ncu-rep --page raw --csv | gzip > gpurun_out/${TAG}_ncu_all_kernels_small_raw.csv.gz
ls -la gpurun_out | tail -15
