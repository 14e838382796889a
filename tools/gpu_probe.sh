#!/bin/bash
# One gpurun call that answers the open measurement questions listed in DESIGN.md §8 (run from the repo root on a B200):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_probe.sh > gpurun_out/probe.txt 2>&1'
# Every line of the output is "<label> <value scans/s> <e2e> <lio ms> <mesh ms> <k_solve_warp ms/scan> <k_pinv ms/scan>".
set -u
line() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --streams "${STREAMS:-1}" 2>/dev/null | grep '^{' | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read()); k = d['kernel_ms_per_scan']
print('$label', d['value'], d['e2e']['value'], d['stage_ms']['lio_total'], d['stage_ms']['mesh_total'], k.get('k_solve_warp'), k.get('k_pinv'), (d.get('multi_stream') or {}).get('value'))"
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 1 3; do line "lu_variant=$v" IMMESH_LU_VARIANT=$v; done                     # REDUX pivot search vs shuffle tournament
for b in 1 2 4; do line "lio_bps=$b" IMMESH_LIO_BPS=$b; done                          # resident blocks per SM of the persistent kernels
for b in 1 2 3; do line "mesh_bps=$b" IMMESH_MESH_BPS=$b; done
STREAMS=4 line "streams=4" IMMESH_LIO_BPS=4 IMMESH_MESH_BPS=3
STREAMS=4 line "streams=4,bps=1" IMMESH_LIO_BPS=1 IMMESH_MESH_BPS=1
STREAMS=8 line "streams=8,bps=1" IMMESH_LIO_BPS=1 IMMESH_MESH_BPS=1
python tools/mini_stream.py 2 2>&1 | tail -3                                          # inverse phase stamps + front-end timing
