#!/bin/bash
# Second capture of round 2 (after the d = 64 softmax change and the measured reference arm): the reference arm and the default bench line
# back to back as the driver runs them, the launch list of the timed region, ncu --set full of the d = 64 attention kernel.
#   gpurun --timeout 1500 -- 'bash profiles/scripts/capture_r02b.sh'
set -x
O=gpurun_out/r02u
mkdir -p $O
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-wan-sp --no-decode > $O/plain_bench.log 2>&1 &&
LTXB200_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-wan-sp --no-decode > $O/ncu_bench.log 2>&1
python profiles/scripts/prof_kernels.py all 20 > $O/prof_kernels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o /tmp/attn64 python profiles/scripts/prof_kernels.py attn 3 > $O/ncu_attn.log 2>&1
ncu -i /tmp/attn64.ncu-rep --page raw --csv > $O/ncu_attn_d64_raw.csv 2>/dev/null
ncu -i /tmp/attn64.ncu-rep --page source --csv > /tmp/attn64_source.csv 2>/dev/null
python profiles/scripts/reduce_source.py /tmp/attn64_source.csv $O/ncu_attn_d64_stalls.csv 40
cut -c1-700 $O/bench_reference.json; cut -c1-1500 $O/bench_final.json; tail -20 $O/prof_kernels.log
