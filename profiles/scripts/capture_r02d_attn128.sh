#!/bin/bash
# ncu --set full of the Wan-1.3B self-attention launch (d = 128, 12 x 32760): the default single-CTA kernel and the opt-in CTA-pair kernel
O=gpurun_out/r02d
mkdir -p $O
for on in 0 1; do
  LTXB200_ATTN128_2CTA=$on timeout 150 ncu --set full --clock-control none --import-source on -k regex:attention -s 3 -c 1 -o /tmp/a128_$on \
      python profiles/scripts/prof_kernels.py attn128 3 > $O/ncu_attn128_$on.log 2>&1
  ncu -i /tmp/a128_$on.ncu-rep --page raw --csv > $O/ncu_attn_d128_2cta${on}_raw.csv 2>/dev/null
  ncu -i /tmp/a128_$on.ncu-rep --page source --csv > /tmp/a128_${on}_source.csv 2>/dev/null
  python profiles/scripts/reduce_source.py /tmp/a128_${on}_source.csv $O/ncu_attn_d128_2cta${on}_stalls.csv 40
  head -3 $O/ncu_attn_d128_2cta${on}_stalls.csv | cut -c1-400
done
