#!/bin/bash
# GPU call 3 of round 2: A/B of kernel tunings, proper ncu captures, failing tests re-run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python scripts/ab_variants.py fd_ ss_ > gpurun_out/c3_ab.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "online or stft_scm or tango or fullsize or filter_dual or z_layout" > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
for t in stft_scm2 stft_scm1 stft stft_scm_c8 stft_scm_c8_256; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_scm_kernel -s 2 -c 1 -o gpurun_out/c3_prof_$t -f python scripts/prof_target.py $t > gpurun_out/c3_ncu_$t.log 2>&1
done
for t in filter_dual:filter_dual masked_scm_zf8:masked_scm filter_multi44:filter_sum_multi tango_mid44:tango_mid tango_mid28:tango_mid solve8:mwf_solve istft:istft; do
  w=${t%%:*}; k=${t##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/c3_prof_$w -f python scripts/prof_target.py $w > gpurun_out/c3_ncu_$w.log 2>&1
done
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/c3_bench_cfg3.json 2> gpurun_out/c3_bench_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/c3_launches_cfg3.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c3_ncu_l3.log 2>&1
cat gpurun_out/c3_ab.log; tail -5 gpurun_out/c3_pytest.log
