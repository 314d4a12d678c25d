#!/bin/bash
nvidia-smi -L | wc -l
python -m pytest tests -m gpu -q -k "real_gpus or several_gpus or across or group" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
python scratch/show_bench.py gpurun_out/r2r_bench_n2.json || strings gpurun_out/r2r_bench_n2.err | tail -20
