#!/bin/bash
# Final capture of round 2 (build with the CTA-pair d = 128 kernel and the diagnostic switches compiled in, all off): the GPU suite, smoke,
# the default bench line, the launch list of the timed region, ncu --set full of the d = 64 attention kernel.
#   gpurun --timeout 1200 -- 'bash profiles/scripts/capture_r02d.sh'
set -x
O=gpurun_out/r02d
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; cut -c1-600 $O/bench_final.json
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-wan-sp --no-decode > $O/plain_bench.log 2>&1 &&
LTXB200_NCU_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-wan-sp --no-decode > $O/ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o /tmp/attn64 python profiles/scripts/prof_kernels.py attn 3 > $O/ncu_attn.log 2>&1
ncu -i /tmp/attn64.ncu-rep --page raw --csv > $O/ncu_attn_d64_raw.csv 2>/dev/null
ncu -i /tmp/attn64.ncu-rep --page source --csv > /tmp/attn64_source.csv 2>/dev/null
python profiles/scripts/reduce_source.py /tmp/attn64_source.csv $O/ncu_attn_d64_stalls.csv 40
ls -la $O
