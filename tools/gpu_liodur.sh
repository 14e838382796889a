#!/bin/bash
# warm-cache durations of the IESKF kernels (ncu single-metric launch list, caches not flushed between kernels) + in-kernel latency stamps
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_liodur.sh r02t "1"'     (second argument: IMMESH_LIO_SPLIT values to run)
tag=${1:-r02x}
modes=${2:-1}
for c in C100k C3 C5; do
  for sp in $modes; do
    IMMESH_LIO_SPLIT=$sp ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv -k regex:"k_match|k_terms|k_residual|k_prepare|k_grow" \
       --log-file gpurun_out/liodur_${tag}_${c}_$sp.csv python tools/debug/lio_only.py $c 16 > gpurun_out/liodur_${tag}_${c}_$sp.log 2>&1
  done
done
if [ -f tools/debug/libimmesh_stamps.so ]; then
  for c in C100k C3; do python tools/debug/lio_stamps.py $c 24 > gpurun_out/stamps_${tag}_$c.txt 2>&1; done
fi
