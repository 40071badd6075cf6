#!/bin/bash
# Re-capture after the cost-kernel change: SGBM launch list + --set full of sgbm_cost_fused (plain run first, each must exit 0).
mkdir -p gpurun_out
tag=${1:-r02d}
base="python bench.py --steps 1 --warmup 3 --min-region-s 0 --no-cpu --no-check --no-latency --workload sgbm720"
$base > gpurun_out/${tag}_plain_sgbm720.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 \
    --csv --log-file gpurun_out/${tag}_launches_sgbm720.csv $base > gpurun_out/${tag}_ncu_sgbm720.log 2>&1
echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgbm_cost_fused -s 2 -c 1 \
    -o gpurun_out/${tag}_prof_sgbm_cost -f $base > gpurun_out/${tag}_ncufull_cost.log 2>&1
echo "full cost rc=$?"
