# same-box A/B of attention builds: LIBS="name name ..." (files ltx-video-gpupoor_b200/<name>.so), default build first and last
R=/root/repo/ltx-video-gpupoor_b200
for lib in libltx_b200 $LIBS libltx_b200; do
echo "== $lib"
REPS=${REPS:-20} LTXB200_LIB=$R/$lib.so timeout 60 python profiles/scripts/attn_ablation_probe.py 2>&1 | grep "TFLOP\|rror" | head -4
done
