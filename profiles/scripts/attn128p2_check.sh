# bring-up of the CTA-pair d = 128 attention kernel: parity tests on the trap-on-hang build, then the regular build, then timing A/B
R=/root/repo/ltx-video-gpupoor_b200
export LTXB200_ATTN128_2CTA=1
echo "== debug build"
LTXB200_LIB=$R/lib_dbg.so timeout 240 python -m pytest tests/test_kernels_gpu.py -k "attention and not 64" -x -q 2>&1 | tail -15
echo "== regular build"
timeout 240 python -m pytest tests/test_kernels_gpu.py -k "attention and not 64" -x -q 2>&1 | tail -5
for on in 0 1 0 1; do
echo "== LTXB200_ATTN128_2CTA=$on"
LTXB200_ATTN128_2CTA=$on REPS=20 timeout 60 python profiles/scripts/attn_ablation_probe.py 2>&1 | grep "TFLOP\|rror" | head -4
done
