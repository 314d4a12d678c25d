#!/bin/bash
# round 2, two GPUs: the multi-GPU tests (group in one process, exchange with device-side flags, fused reduce, rt_headless --gpus),
# then bench.py at N=2 (weak headline + strong leg + in-bench check)
nvidia-smi -L
python -m pytest tests -m gpu -q -k "real_gpus or several_gpus or across" -s 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -c 1500 gpurun_out/r2f_bench_n2.err
python scratch/show_bench.py gpurun_out/r2f_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --reduce nccl > gpurun_out/r2f_bench_n2_nccl.json 2> gpurun_out/r2f_bench_n2_nccl.err
python scratch/show_bench.py gpurun_out/r2f_bench_n2_nccl.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 --config c3 --scaling strong --spp 32 > gpurun_out/r2f_bench_c3_strong_n2.json 2> gpurun_out/r2f_bench_c3_strong_n2.err
python scratch/show_bench.py gpurun_out/r2f_bench_c3_strong_n2.json
