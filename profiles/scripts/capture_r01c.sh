#!/bin/bash
# Third capture of round 1 (after the CTA-pair GEMM and the pipelined epilogue): bench lines, launch list of the timed region,
# `ncu --set full` of the three LTX GEMM shapes (reduced to CSV on the box).
set -x
O=gpurun_out
T=r01c
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_final.json 2> $O/${T}_bench_final.err
python bench.py --steps 5 --warmup 3 --workload ltx2b_768x512x121_i2v_cfg_stg --no-cpu-baseline > $O/${T}_bench_ltx_i2v.json 2> $O/${T}_bench_ltx_i2v.err
python bench.py --steps 3 --warmup 3 --workload wan1.3b_832x480x81_sp --no-cpu-baseline > $O/${T}_bench_wan1gpu.json 2> $O/${T}_bench_wan1gpu.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > $O/plain_bench.log 2>&1 &&
LTXB200_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > $O/ncu_bench.log 2>&1
python profiles/scripts/prof_kernels.py all 20 > $O/${T}_prof_kernels.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 7 -o /tmp/gemm python profiles/scripts/prof_kernels.py gemm 1 > $O/ncu_gemm.log 2>&1
ncu -i /tmp/gemm.ncu-rep --page raw --csv > $O/${T}_gemm_raw.csv 2>/dev/null
cat $O/${T}_prof_kernels.log
