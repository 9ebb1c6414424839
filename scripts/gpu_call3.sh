#!/bin/bash
# GPU call of round 2: A/B of kernel tunings, ncu captures (exported to CSV on the box: gpurun_out must stay < 64 MiB)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python scripts/ab_variants.py fd_ ss_ > gpurun_out/c3_ab.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "online or stft_scm or tango or fullsize or filter_dual or z_layout or dnn" > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
prof() {   # prof <target> <kernel regex>
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/p_$1 -f python scripts/prof_target.py $1 > gpurun_out/c3_ncu_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/c3_raw_$1.csv 2>/dev/null
  ncu -i /tmp/p_$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/c3_src_$1.csv.gz
}
prof stft_scm2 stft_scm_kernel
cp /tmp/p_stft_scm2.ncu-rep gpurun_out/c3_prof_stft_scm2.ncu-rep
prof stft_scm1 stft_scm_kernel
prof stft stft_scm_kernel
prof stft_scm_c8 stft_scm_kernel
prof stft_scm_c8_256 stft_scm_kernel
prof filter_dual filter_dual
prof masked_scm_zf8 masked_scm
prof filter_multi44 filter_sum_multi
prof tango_mid44 tango_mid
prof tango_mid28 tango_mid
prof solve8 mwf_solve
prof istft istft
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/c3_bench_cfg3.json 2> gpurun_out/c3_bench_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/c3_launches_cfg3.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c3_ncu_l3.log 2>&1
du -sh gpurun_out; cat gpurun_out/c3_ab.log; tail -5 gpurun_out/c3_pytest.log
