# same-box A/B of attention builds under sustained (power-capped) conditions: LIBS="name ..." (ltx-video-gpupoor_b200/<name>.so)
R=/root/repo/ltx-video-gpupoor_b200
for lib in libltx_b200 $LIBS libltx_b200; do
echo "== $lib"
LTXB200_LIB=$R/$lib.so timeout 60 python profiles/scripts/sustained_attn_only.py 2>&1 | grep "sustained\|rror" | head -4
done
