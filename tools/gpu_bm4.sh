#!/bin/bash
# bm_sad4 iteration call: parity check against bm_sad3, stage timing, optionally one ncu --set full capture of the kernel
mkdir -p gpurun_out
tag=$1
timeout 150 python tools/bm4_check.py ${2:-all} > gpurun_out/${tag}_check.log 2>&1; echo "check rc=$?"; grep -v "diff 0 raw diff 0 cost diff 0 final diff 0" gpurun_out/${tag}_check.log | tail -25
if [ "$3" = "ncu" ]; then
  cmd="python bench.py --workload bm720 --steps 1 --warmup 3 --min-region-s 0 --no-cpu --no-check"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm_sad4 -s 2 -c 1 -o gpurun_out/${tag}_prof_bm4 -f $cmd > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
fi
