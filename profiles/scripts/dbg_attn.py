import sys, os, torch
sys.path.insert(0, os.getcwd())
from ltx_video_gpupoor_b200 import ops
torch.manual_seed(0)
B,H,Lq,Lk,d = 1,2,128,128,64
q,k,v=[torch.randn(B,L,H,d,device='cuda').bfloat16() for L in (Lq,Lk,Lk)]
o=ops.attention(q,k,v); torch.cuda.synchronize()
ref=torch.nn.functional.scaled_dot_product_attention(q.transpose(1,2).float(),k.transpose(1,2).float(),v.transpose(1,2).float()).transpose(1,2)
print('err', ((o.float()-ref).norm()/ref.norm()).item())
