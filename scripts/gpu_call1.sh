#!/bin/bash
# GPU call 1 of round 2: first contact of the new kernels, tests, first bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python scripts/quick_check.py > gpurun_out/c1_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/c1_quick.log
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/c1_bench_cfg2.json 2> gpurun_out/c1_bench_cfg2.err
timeout 300 python bench.py --workload cfg4_512 --steps 20 --warmup 3 --no-cpu > gpurun_out/c1_bench_cfg4_512.json 2> gpurun_out/c1_bench_cfg4_512.err
tail -3 gpurun_out/c1_quick.log; tail -3 gpurun_out/c1_pytest.log; head -c 600 gpurun_out/c1_bench_cfg2.json
