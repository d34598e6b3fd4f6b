#!/bin/bash
# Compile every experimental variant of libltx_b200.so (compile-time switches of DESIGN.md §9) into /tmp: a "do they still build" check
# (no GPU needed); run the GPU suite against one with LTXB200_LIB=/tmp/lib_<name>.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/../../ltx-video-gpupoor_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $F $2 -o /tmp/lib_$1.so ltx_b200.cu && echo "ok   $1 ($2)" || echo "FAIL $1 ($2)"; }
build ones "-DLTXB200_ATTN64_ONES" &
build half64 "-DLTXB200_ATTN64_HALFROW" &
build half128 "-DLTXB200_ATTN128_HALFROW" &
build bn64 "-DLTXB200_ATTN128_BN64" &
wait
build split34 "-DLTXB200_ATTN128_SPLIT34" &
build mma2 "-DLTXB200_ATTN128_2MMA" &
build pingpong "-DLTXB200_ATTN128_PINGPONG" &
build gelu_scalar "-DLTXB200_GELU_SCALAR" &
wait
build poly "-DLTXB200_ATTN_POLY_D64=2 -DLTXB200_ATTN_POLY_DEG_D64=3 -DLTXB200_ATTN_POLY_D128=0" &
build loadall "-DLTXB200_ATTN_LOADALL=2" &
build l2pf "-DLTXB200_ATTN_L2PF=4 -DLTXB200_ATTN_PRODUCER_PARK" &
build ablations "-DLTXB200_ABL_NOEXP -DLTXB200_ABL_NOMAX -DLTXB200_ABL_NOSUM -DLTXB200_ABL_NOLOAD -DLTXB200_ABL_NOTMEM" &
build freemma "-DLTXB200_ABL_FREEMMA" &
build dbg "-DLTXB200_DEBUG_HANG" &
wait
