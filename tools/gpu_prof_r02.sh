#!/bin/bash
# Round-2 profile call: launch lists of both workloads, --set full captures of bm_sad4 and of every SGBM kernel.
mkdir -p gpurun_out
tag=${1:-r02}
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
full bm4 bm720 'bm_sad4' 2 1
full bmsmall bm720 'prefilter|post_row8|speckle|morph' 0 7
full sgbm_small sgbm720 'sgbm_planes2|sgbm_cost_fused|sgbm_path4|sgbm_lr|median3' 0 6
full sgbm_sweep sgbm720 'sgbm_sweep' 100 1
