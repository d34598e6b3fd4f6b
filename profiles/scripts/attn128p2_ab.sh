# CTA-pair d = 128 kernel: parity, then same-box timing A/B against the single-CTA kernel
LTXB200_ATTN128_2CTA=1 timeout 120 python profiles/scripts/attn128p2_check.py 2>&1 | tail -3
for on in 0 1 0 1; do
echo "== LTXB200_ATTN128_2CTA=$on"
LTXB200_ATTN128_2CTA=$on REPS=20 timeout 60 python profiles/scripts/attn_ablation_probe.py 2>&1 | grep "d128\|rror" | head -4
done
