#!/bin/bash
# Round-2 FINAL-build profile capture (one gpurun call): launch list of ~4 training steps and per-launch section metrics of the
# tensor-core kernels of one step.  The same bench command runs once WITHOUT ncu first.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_BENCH_EXTRA_WARMUP=0
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-sample --no-cudnn --no-swap"
$B > gpurun_out/prof_r02c_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_r02c_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1700 --csv \
  --log-file gpurun_out/launches_r02c.csv $B > gpurun_out/prof_r02c_ncu1.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_r02c.csv)"
timeout 1200 ncu --section SpeedOfLight --section SchedulerStats --section Occupancy --section LaunchStats --clock-control none \
  -k regex:"conv_tc_kernel|conv_slab_kernel|wgrad_tc_kernel|wgrad_slab_kernel|head_conv_kernel" -s 600 -c 160 --csv --page raw \
  --log-file gpurun_out/conv_sections_r02c.csv $B > gpurun_out/prof_r02c_ncu2.log 2>&1
echo "sections rc=$? lines=$(wc -l < gpurun_out/conv_sections_r02c.csv)"
gzip -9 -f gpurun_out/launches_r02c.csv
