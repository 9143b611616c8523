#!/bin/bash
# ncu --set full captures of single ops (one launch each).  The .ncu-rep files (17 MB each with the source import) are
# exported to text on the box and removed: gpurun brings back at most 64 MiB.
cap() {  # name kind M N K mode
  local rep=gpurun_out/$1.ncu-rep
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -c ${NCU_COUNT:-1} -f -o gpurun_out/$1 \
    python tools/one_op.py $2 $3 $4 $5 $6 > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log
  [ -f $rep ] || return
  ncu -i $rep --page details > gpurun_out/$1_details.txt 2>&1
  ncu -i $rep --page raw --csv > gpurun_out/$1_raw.csv 2>&1
  ncu -i $rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1_source.csv.gz
  rm -f $rep
}
for spec in "$@"; do
  cap $spec
done
