#!/bin/bash
# One gpurun call: GPU tests, the default bench line, stream-priority experiment, the other BASELINE configs.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r02f "C1 C2 C3 C5"'
tag=${1:-r02x}; cfgs=${2:-"C3"}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_$tag.txt
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
for p in ${PRIO_EXPERIMENTS:-}; do
  env $p timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --streams 1 --no-raw-leg > gpurun_out/bench_${tag}_${p%%=*}.json 2>/dev/null
done
for c in $cfgs; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err
  tail -2 gpurun_out/bench_${tag}_$c.err
done
