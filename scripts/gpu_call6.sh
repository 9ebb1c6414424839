#!/bin/bash
# NOT RUN in round 2: the 8-GPU call before it used up the round's GPU minutes (two bench processes hung in the NCCL
# teardown after CUDA-graph replays until their timeouts; bench.py now leaves through os._exit in that mode).
# Final single-GPU call of round 2: whole suite on the final kernels, bench lines of every BASELINE shape, CRNN e2e
# variants, launch lists, ncu captures of the kernels that changed last (exported to CSV on the box)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c6_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c6_smoke.log
timeout 600 python scripts/quick_check.py > gpurun_out/c6_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/c6_quick.log
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/c6_bench_cfg2.json 2> gpurun_out/c6_bench_cfg2.err
timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu --masks crnn > gpurun_out/c6_bench_cfg2_crnn.json 2> gpurun_out/c6_bench_cfg2_crnn.err
timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu --masks crnn --crnn-bf16 > gpurun_out/c6_bench_cfg2_crnn_bf16.json 2> gpurun_out/c6_bench_cfg2_crnn_bf16.err
for w in cfg3 cfg5 cfg4_512 cfg4_256 cfg4_1024; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/c6_bench_$w.json 2> gpurun_out/c6_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/c6_bench_reference.json 2> gpurun_out/c6_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/c6_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/c6_ncu_l2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file gpurun_out/c6_launches_cfg4_512.csv python bench.py --workload cfg4_512 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/c6_ncu_l4.log 2>&1
prof() {   # prof <target> <kernel regex>
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/p_$1 -f python scripts/prof_target.py $1 > gpurun_out/c6_ncu_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/c6_raw_$1.csv 2>/dev/null
  ncu -i /tmp/p_$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/c6_src_$1.csv.gz
}
prof stft_scm2 stft_scm_kernel
cp /tmp/p_stft_scm2.ncu-rep gpurun_out/c6_prof_stft_scm2.ncu-rep
prof stft_scm1 stft_scm_kernel
prof stft_scm_c8 stft_scm_kernel
prof stft_scm_c8_256 stft_scm_kernel
du -sh gpurun_out; tail -4 gpurun_out/c6_pytest.log; tail -2 gpurun_out/c6_smoke.log; tail -2 gpurun_out/c6_quick.log; head -c 300 gpurun_out/c6_bench_cfg2.json
