# same-box A/B of the d = 64 self-attention kernels: attention_fwd_kernel<64,false> (LTXB200_ATTN64P=0) vs attention64p_kernel
mkdir -p gpurun_out/r02s
timeout 150 python profiles/scripts/attn64p_check.py 2>&1 | grep -v "^B" | head -3
for i in 1 2; do
LTXB200_ATTN64P=0 timeout 150 python profiles/scripts/prof_kernels.py attn 30 2>&1 | grep "attention d64 B3 N6144 H32"
LTXB200_ATTN64P=1 timeout 150 python profiles/scripts/prof_kernels.py attn 30 2>&1 | grep "attention d64 B3 N6144 H32"
done
