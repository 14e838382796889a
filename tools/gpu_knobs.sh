#!/bin/bash
# knob sweep on one GPU: python bench lines with different IMMESH_* settings (value, e2e, lio ms, mesh ms, tri_warp, mesh256, dilate)
line() { local label=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --streams 1 --no-raw-leg 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_scan']
print('$label', d['value'], d['e2e']['value'], d['stage_ms']['lio_total'], d['stage_ms']['mesh_total'], k.get('k_voxel_tri_warp'), k.get('k_voxel_mesh<256>'), k.get('k_voxel_dilate'), (d.get('parity_gate') or {}).get('ok'))"; }
line base X=1
line nmax64 IMMESH_WARP_NMAX=64
line nmax48 IMMESH_WARP_NMAX=48
line nmax32 IMMESH_WARP_NMAX=32
line dil4 IMMESH_DILATE_BPS=4
line nmax48_dil4 IMMESH_WARP_NMAX=48 IMMESH_DILATE_BPS=4
line nmax48_dil4_bps4 IMMESH_WARP_NMAX=48 IMMESH_DILATE_BPS=4 IMMESH_MESH_BPS=4
