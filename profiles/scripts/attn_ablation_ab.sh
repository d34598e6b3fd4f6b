R=/root/repo/ltx-video-gpupoor_b200
for lib in libltx_b200 abl_base abl_noload abl_notmem abl_freemma abl_freemma_noload libltx_b200; do
echo "== $lib"
LTXB200_LIB=$R/$lib.so timeout 60 python profiles/scripts/attn_ablation_probe.py 2>&1 | grep "TFLOP\|rror" | head -4
done
