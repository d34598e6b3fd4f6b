#!/bin/bash
# Round-1 evidence capture (run under gpurun on one B200): bench lines, the ncu launch list of the timed region, and one
# `ncu --set full` capture per hot kernel, reduced on the box to CSV (the .ncu-rep files exceed gpurun's 64 MiB return limit).
set -x
O=gpurun_out
T=${TAG:-r01}     # file prefix: r01 = first capture of the round, r01b = after the packed/poly softmax and the rope prefetch
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_final.json 2> $O/${T}_bench_final.err
python bench.py --steps 3 --warmup 3 --workload wan1.3b_832x480x81_sp > $O/${T}_bench_wan1gpu.json 2> $O/${T}_bench_wan1gpu.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > $O/plain_bench.log 2>&1 &&
LTXB200_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > $O/ncu_bench.log 2>&1
python profiles/scripts/prof_kernels.py all 1 > $O/plain_all.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip, count, script arg, with_source
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/$1 python profiles/scripts/prof_kernels.py $5 1 > $O/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/${T}_$1_raw.csv 2>/dev/null
  if [ "$6" = "src" ]; then ncu -i /tmp/$1.ncu-rep --page source --csv > $O/${T}_$1_source.csv 2>/dev/null; fi
  rm -f /tmp/$1.ncu-rep
}
cap attn_d64 attention_fwd 2 1 attn src          # 3rd launch: self-attention B3 N6144 H32 d64
cap xattn_d64 attention_fwd 5 1 attn nosrc       # 6th launch: cross-attention with bias
cap attn_d128 attention_fwd 2 1 attn128 src      # Wan-1.3B shape
cap gemm gemm_bf16 2 7 gemm nosrc                # ffn_up (3rd), ffn_down (6th), qkv (9th) are launches 2,5,8
cap ew "norm_mod|qk_norm_rope" 3 9 ew nosrc
ls -la $O | tail -25
