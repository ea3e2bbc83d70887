#!/bin/bash
# Run on the GPU box (via gpurun): launch lists + one full ncu capture per dominant kernel.
# usage: benchmarks/profile_round.sh <tag>     outputs -> gpurun_out/  (summarise with ncu_summary.py)
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_icm_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_icm_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icm_ils -c 1 -o gpurun_out/icm_${TAG} \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 >> gpurun_out/ncu_icm_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_adc_${TAG}.csv \
    python bench.py --workload adc --m 8 --reps 1 --cpu-queries 2 --check-queries 2 > gpurun_out/ncu_adc_${TAG}.log 2>&1
# the tensor-core ADC path: sample-values / sample-lists / main filter are launches 0, 1, 2 of adc_filter_kernel per call
# (two warm-up calls first), plus the kernels around it; LSQ_B200_ADC=scan for the lookup kernel
ncu --set full --clock-control none --import-source on -k regex:"adc_filter_kernel|adc_decode_kernel|adc_rescore_kernel|adc_lut_rows_kernel|adc_sample_tau_kernel|threshold_kernel" -s 16 -c 8 -o gpurun_out/adc_tc_${TAG} \
    python bench.py --workload adc --m 16 --reps 1 --cpu-queries 2 --check-queries 2 >> gpurun_out/ncu_adc_${TAG}.log 2>&1
LSQ_B200_ADC=scan ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/scan_${TAG} \
    python bench.py --workload adc --m 8 --reps 1 --cpu-queries 2 --check-queries 2 >> gpurun_out/ncu_adc_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi -s 1 -c 1 -o gpurun_out/viterbi_${TAG} \
    python bench.py --workload chain --m 8 --n 400000 --reps 1 --cpu-n 10 > gpurun_out/ncu_vit_${TAG}.log 2>&1
