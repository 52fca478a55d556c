#!/bin/bash
# first GPU call of round 2: full GPU test suite, bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt; lscpu | head -20 >> gpurun_out/c1_nproc.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
echo "bench ref rc=$?"
tail -c 600 gpurun_out/c1_bench.json
