import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from oracle import ltx_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
z = torch.randn(1, 128, 2, 4, 6, generator=torch.Generator().manual_seed(2))
sd = O.make_vae_decoder_state_dict(seed=1)
for cudnn in (True, False):
    torch.backends.cudnn.enabled = cudnn
    ca, cb = [], []
    with torch.no_grad():
        O.vae_decode(sd, z, collect=ca)
        O.vae_decode({k: v.cuda() for k, v in sd.items()}, z.cuda(), collect=cb)
    for i, (a, b) in enumerate(zip(ca, cb)):
        print(f"cudnn={cudnn} stage {i} shape {tuple(a.shape)}: gpu vs cpu rel_l2 = {O.rel_l2(b.cpu(), a):.2e}", flush=True)
# the suspicious op in isolation: replicate-padded conv on a tiny grid, many output channels
torch.backends.cudnn.enabled = True
g = torch.Generator().manual_seed(0)
for (cin, cout, T, H, W) in ((128, 512, 4, 4, 6), (512, 512, 4, 4, 6), (512, 4096, 4, 4, 6), (512, 2048, 4, 4, 6), (256, 1024, 4, 8, 12)):
    x = torch.randn(1, cin, T, H, W, generator=g); w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5; b = torch.randn(cout, generator=g)
    ref = F.conv3d(x, w, b, padding=(0, 1, 1))
    y = F.conv3d(x.cuda(), w.cuda(), b.cuda(), padding=(0, 1, 1)).cpu()
    print(f"conv3d {cin}->{cout} {T}x{H}x{W}: gpu vs cpu rel_l2 = {O.rel_l2(y, ref):.2e}")
