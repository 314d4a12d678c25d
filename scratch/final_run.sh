set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r1B_bench_n1.json 2> gpurun_out/r1B_bench_n1.err; tail -c 600 gpurun_out/r1B_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1B_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/r1B_bench_reference_arm.json
rm -f gpurun_out/r1B_bench_configs.jsonl
for cfg in "c3 --spp 16" "c4 --spp 64" "c4 --spp 256" "c5" "c1"; do python bench.py --config $cfg --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r1B_bench_configs.jsonl; done
cut -c1-220 gpurun_out/r1B_bench_configs.jsonl
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
