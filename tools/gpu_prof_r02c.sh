#!/bin/bash
# Round-2 final profile call: launch lists of both workloads (BM with bm_sad4, SGBM with the whole-height cluster passes) and
# --set full captures of the SGBM kernels that changed (sgbm_vpass, sgbm_path4 first / last).  Each ncu run follows a plain run
# of the same command that exited 0.
mkdir -p gpurun_out
tag=${1:-r02c}
base="python bench.py --steps 1 --warmup 3 --min-region-s 0 --no-cpu --no-check --no-latency"
for wl in bm720 sgbm720; do
  cmd="$base --workload $wl"
  $cmd > gpurun_out/${tag}_plain_$wl.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 \
      --csv --log-file gpurun_out/${tag}_launches_$wl.csv $cmd > gpurun_out/${tag}_ncu_$wl.log 2>&1
  echo "launches $wl rc=$?"
done
full() {  # name workload regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 \
      -o gpurun_out/${tag}_prof_$1 -f $base --workload $2 > gpurun_out/${tag}_ncufull_$1.log 2>&1
  echo "full $1 rc=$?"
}
full sgbm_vpass sgbm720 'sgbm_vpass' 2 1
full sgbm_paths sgbm720 'sgbm_path4|sgbm_cost_fused' 3 3
