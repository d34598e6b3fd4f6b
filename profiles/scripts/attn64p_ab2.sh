# same-box A/B of softmax-polynomial variants (degree, share of the exponentials) x {attention_fwd_kernel<64>, attention64p_kernel}
R=/root/repo/ltx-video-gpupoor_b200
for lib in ${LIBS:-libltx_b200.so}; do
for p in 0 1; do
echo "== $lib ATTN64P=$p"
LTXB200_LIB=$R/$lib LTXB200_ATTN64P=$p timeout 150 python profiles/scripts/attn64p_check.py 2>&1 | grep "worst\|TFLOP" | head -3
done
done
