#!/bin/bash
# Multi-GPU call (gpurun --gpus N): topology, concurrent host-link ceiling, row-band split over peer copies, bench at N ranks.
N=${1:-8}
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/multi_topo.txt 2>&1
timeout 200 python tools/pcie_concurrent.py $N > gpurun_out/multi_pcie.csv 2> gpurun_out/multi_pcie.err; echo "pcie rc=$?"; cat gpurun_out/multi_pcie.csv
timeout 200 python tools/pcie_concurrent.py $N --affinity > gpurun_out/multi_pcie_aff.csv 2>> gpurun_out/multi_pcie.err; tail -n +2 gpurun_out/multi_pcie_aff.csv
timeout 200 python -m pytest tests/test_rowband_gpu.py -x -q -m gpu -k all_gpus 2>&1 | tail -2
timeout 300 python tools/rowband_bench.py $N > gpurun_out/multi_rowband.json 2> gpurun_out/multi_rowband.err; echo "rowband rc=$?"; cat gpurun_out/multi_rowband.json
for k in 2 4 8; do
  [ $k -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $k --steps 20 --warmup 3 --workload bm720 --no-cpu \
      > gpurun_out/multi_bench_$k.json 2> gpurun_out/multi_bench_$k.err; echo "bench $k rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/multi_bench_$k.json").read().strip().splitlines()[-1])
    e = d["e2e"]; print("N=$k value", round(d["value"]), "fps", round(d["fps"]), "e2e fps", round(e["fps"]), "link", round(e["link_gbs"], 1), "/", round(e["link_ceiling_gbs"], 1), "frac", round(e["link_frac"], 3))
except Exception as ex: print("parse failed", ex)
PY
done
