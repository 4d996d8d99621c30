import ctypes as C, sys, torch
sys.path.insert(0, '/root/repo')
from s3prl_b200 import lib
L = lib.load()
B,T,H = 32,499,12
D = H*64
q = torch.randn(B,T,D,device='cuda'); k = torch.randn(B,T,D,device='cuda'); v = torch.randn(B,T,D,device='cuda')
out = torch.empty(B,T,D,device='cuda')
vf = (C.c_int32*B)(*([T]*B))
for i in range(2):
    lib.check(L.s3b_attention_f32(C.c_void_p(q.data_ptr()),C.c_void_p(k.data_ptr()),C.c_void_p(v.data_ptr()),vf,B,T,H,C.c_void_p(out.data_ptr()),None))
torch.cuda.synchronize()
