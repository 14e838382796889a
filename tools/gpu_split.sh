#!/bin/bash
# A/B of the 8-lanes-per-point match (k_match + k_terms, default) against the one-thread-per-point residual pass (IMMESH_LIO_SPLIT=0).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_split.sh r02p'
tag=${1:-r02x}
python -m pytest tests/test_parity_gpu.py tests/test_lio_gpu.py -m gpu -x -q -k "not torchrun and not sharded" 2>&1 | tail -6 > gpurun_out/t_$tag.txt
for c in C100k C3 C5; do
  timeout 300 python bench.py --config $c --no-cpu-baseline --no-raw-leg --streams 1 > gpurun_out/bench_${tag}_${c}_split.json 2> gpurun_out/bench_${tag}_${c}_split.err
  IMMESH_LIO_SPLIT=0 timeout 300 python bench.py --config $c --no-cpu-baseline --no-raw-leg --streams 1 > gpurun_out/bench_${tag}_${c}_serial.json 2> gpurun_out/bench_${tag}_${c}_serial.err
done
