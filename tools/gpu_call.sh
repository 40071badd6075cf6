#!/bin/bash
# One gpurun call: GPU tests, the default bench line, then ncu launch lists and --set full captures of the small kernels.
# usage (from the repo root, on the GPU box):  bash tools/gpu_call.sh <tag> [tests] [bench] [launches] [full:<workload>:<regex>:<skip>:<count>] ...
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err;;
    launches)
      for wl in bm720 sgbm720; do
        cmd="python bench.py --workload $wl --steps 1 --warmup 3 --min-region-s 0 --no-cpu --no-check"
        $cmd > gpurun_out/${tag}_plain_$wl.log 2>&1 &&
        timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 \
            --csv --log-file gpurun_out/${tag}_launches_$wl.csv $cmd > gpurun_out/${tag}_ncu_$wl.log 2>&1
        echo "launches $wl rc=$?"
      done;;
    full:*)
      IFS=: read -r _ wl regex skip count <<< "$what"
      cmd="python bench.py --workload $wl --steps 1 --warmup 3 --min-region-s 0 --no-cpu --no-check"
      name=$(echo "$regex" | tr -c 'A-Za-z0-9_\n' '_')
      $cmd > gpurun_out/${tag}_plain_full.log 2>&1 &&
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count \
          -o gpurun_out/${tag}_prof_$name -f $cmd > gpurun_out/${tag}_ncufull_$name.log 2>&1
      echo "full $regex rc=$?";;
  esac
done
