#!/bin/bash
# GPU call 2 of round 2: full checks + timings, whole GPU suite, bench lines, ncu captures of the new kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python scripts/quick_check.py > gpurun_out/c2_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/c2_quick.log
timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/c2_bench_cfg2.json 2> gpurun_out/c2_bench_cfg2.err
for w in cfg4_512 cfg4_256 cfg3 cfg5; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_bench_$w.json 2> gpurun_out/c2_bench_$w.err
done
# ncu: launch list of a short cfg2 run, then full captures of the new kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c2_launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stft_scm_kernel -s 6 -c 1 -o gpurun_out/c2_prof_stft_scm2 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:filter_dual -s 6 -c 1 -o gpurun_out/c2_prof_filter_dual -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stft_scm_kernel -s 4 -c 1 -o gpurun_out/c2_prof_stft_scm_c8 -f \
    python bench.py --workload cfg4_512 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_ncu_c.log 2>&1
tail -3 gpurun_out/c2_quick.log; tail -5 gpurun_out/c2_pytest.log; ls -la gpurun_out | grep c2_
