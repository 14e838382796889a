#!/bin/bash
# One 8-GPU gpurun call: sharded parity at 2/4/8 ranks, C5 sharded over 8 GPUs, C4 sharded over 4, the default bench at N=8.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_round8.sh r02k'
tag=${1:-r02x}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k sharded 2>&1 | tail -5 > gpurun_out/t8_$tag.txt
run 8 29611 --config C5 --mode sharded --no-raw-leg > gpurun_out/bench8_${tag}_C5s.json 2> gpurun_out/bench8_${tag}_C5s.err
run 4 29612 --config C5 --mode sharded --no-raw-leg > gpurun_out/bench4_${tag}_C5s.json 2> gpurun_out/bench4_${tag}_C5s.err
run 4 29613 --config C4 --mode sharded --no-raw-leg > gpurun_out/bench4_${tag}_C4s.json 2> gpurun_out/bench4_${tag}_C4s.err
run 8 29614 --steps 30 --warmup 5 > gpurun_out/bench8_${tag}.json 2> gpurun_out/bench8_${tag}.err
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
