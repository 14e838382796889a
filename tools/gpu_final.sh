#!/bin/bash
# Final single-GPU sweep of a round: GPU tests, the default bench line, every BASELINE config, ncu launch list + full capture.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_final.sh r02n'
tag=${1:-r02x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_$tag.txt
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
for c in C1 C2 C3 C4 C5; do
  timeout 400 python bench.py --config $c > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err
done
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/benchref_$tag.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/mini_stream.py 6 > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_voxel_dilate|k_voxel_tri_warp|k_pull_vertices|k_cand_init|k_grow_voxel|k_grow_simple|k_match|k_terms|k_push_add|k_commit_faces" -s 33 -c 19 -o gpurun_out/prof_$tag python tools/mini_stream.py 5 > gpurun_out/ncu_f_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
if [ -f tools/debug/libimmesh_stamps.so ]; then
  for c in C100k C3; do python tools/debug/lio_stamps.py $c 24 > gpurun_out/stamps_${tag}_$c.txt 2>&1; done
  python tools/debug/mesh_stamps.py C100k 16 > gpurun_out/mstamps_$tag.txt 2>&1
fi
