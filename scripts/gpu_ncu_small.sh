#!/bin/bash
# first invocation of every kernel of the library at reduced sizes, full metric set; the report is turned into the
# per-kernel table on the box (the .ncu-rep itself is larger than what travels back)
mkdir -p gpurun_out
python scripts/run_all_kernels.py small > gpurun_out/plain_all.log 2>&1 &&
ncu --set full --clock-control none --kernel-id :::1 -o /tmp/prof_all_small_r02 -f python scripts/run_all_kernels.py small > gpurun_out/ncu_all.log 2>&1
tail -n 2 gpurun_out/plain_all.log gpurun_out/ncu_all.log
python scripts/ncu_table.py /tmp/prof_all_small_r02.ncu-rep gpurun_out/r02_ncu_all_kernels_small > /dev/null
ncu -i /tmp/prof_all_small_r02.ncu-rep --page raw --csv | gzip > gpurun_out/r02_ncu_all_kernels_small_raw.csv.gz
ls -la gpurun_out
