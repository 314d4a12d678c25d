#!/bin/bash
# r3e N GPUs (N = $1): bench.py weak headline + strong leg + in-bench check
n=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r3e_bench_n$n.json 2> gpurun_out/r3e_bench_n$n.err
python scratch/show_bench.py gpurun_out/r3e_bench_n$n.json | head -3 || strings gpurun_out/r3e_bench_n$n.err | tail -20
