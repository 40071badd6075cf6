#!/bin/bash
# Development call for the SGBM stage: parity tests of the matcher, device-resident timings, optionally one ncu capture.
# usage: bash tools/gpu_sgbm_dev.sh [quick] [ncu] [batch sizes ...]
mkdir -p gpurun_out
sel=(); if [ "$1" = quick ]; then sel=(-k "cluster_pass or 720p or wave_sized"); shift; fi
ncu=0; if [ "$1" = ncu ]; then ncu=1; shift; fi
timeout 900 python -m pytest tests/test_sgbm_gpu.py -x -q -m gpu "${sel[@]}" > gpurun_out/sgbm_dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/sgbm_dev_pytest.log
for b in ${@:-45 15}; do
  timeout 300 python tools/quick_time_sgbm.py $b 2>&1 | tail -2
done
RTDM_SGBM_VPASS_MIN=1 timeout 300 python tools/quick_time_sgbm.py 1 2>&1 | tail -2
if [ $ncu = 1 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgbm_vpass -s 2 -c 1 -o gpurun_out/dev_prof_vpass -f \
      python tools/quick_time_sgbm.py 15 > gpurun_out/dev_ncu_vpass.log 2>&1; echo "ncu rc=$?"
fi
