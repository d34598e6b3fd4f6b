"""Is torch's own conv3d on this box a usable fp32 reference?  (diag for tests/test_zz_full_size_gpu.py)"""
import torch
import torch.nn.functional as F
g = torch.Generator().manual_seed(0)
for (cin, cout, T, H, W) in ((128, 128, 6, 16, 24), (512, 512, 4, 8, 12), (128, 48, 9, 32, 48)):
    x = torch.randn(1, cin, T, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=(0, 1, 1))
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for cudnn in (True, False):
            torch.backends.cudnn.enabled = cudnn
            for dt in (torch.float32, torch.float64, torch.bfloat16):
                y = F.conv3d(x.cuda().to(dt), w.cuda().to(dt), b.cuda().to(dt), padding=(0, 1, 1)).double().cpu()
                e = float((y - ref).norm() / ref.norm())
                print(f"conv3d {cin}->{cout} {T}x{H}x{W} tf32={tf32} cudnn={cudnn} {str(dt)[6:]}: rel_l2 vs cpu fp64 = {e:.2e}", flush=True)
print(torch.__version__, torch.backends.cudnn.version())
