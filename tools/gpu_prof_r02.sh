#!/bin/bash
# Round-2 profile capture (one gpurun call): the launch list of the training step and per-launch section metrics of the
# tensor-core kernels.  The same bench command runs once WITHOUT ncu first.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_BENCH_EXTRA_WARMUP=0
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-sample --no-cudnn --no-swap"
$B > gpurun_out/prof_r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_r02_plain.log; exit 1; }
grep -o '"ms_per_step": [0-9.]*' gpurun_out/prof_r02_plain.log | head -1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_r02.csv $B > gpurun_out/prof_r02_ncu1.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_r02.csv)"
timeout 1500 ncu --section SpeedOfLight --section SchedulerStats --section Occupancy --section LaunchStats --clock-control none \
  -k regex:"conv_tc_kernel|conv_slab_kernel|wgrad_tc_kernel|wgrad_slab_kernel|head_conv_kernel" -s 600 -c 160 --csv --page raw \
  --log-file gpurun_out/conv_sections_r02.csv $B > gpurun_out/prof_r02_ncu2.log 2>&1
echo "sections rc=$? lines=$(wc -l < gpurun_out/conv_sections_r02.csv)"
gzip -9 -f gpurun_out/launches_r02.csv
ls -la gpurun_out/launches_r02.csv.gz gpurun_out/conv_sections_r02.csv
