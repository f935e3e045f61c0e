import sys, os, ctypes, torch
sys.path.insert(0, '/root/repo')
from uda_clr_b200 import _lib, synth
from uda_clr_b200._lib import check, ptr
lib=_lib.load(); dev='cuda'
B,C,H,K=8,256,128,2; HW=H*H
g=torch.Generator().manual_seed(1)
y=synth.nested_ellipse_labels(B,K,H,H,g).to(dev)
xs=[torch.randn(B,C,H,H,device=dev) for _ in range(3)]
D=torch.randn(K,C,device=dev)*0.1; beta=torch.zeros(K,device=dev)
ws_bytes=lib.clr_disc_fused_ws_bytes(C,K); ws=torch.empty(ws_bytes,dtype=torch.uint8,device=dev)
coef=torch.empty(B,K,H,H,device=dev); packed2=torch.empty(K*(C+1)+4,device=dev)
st=torch.cuda.current_stream().cuda_stream
def run(i):
    check(lib.clr_disc_fused_fwd(ptr(xs[i%3]),ptr(y),B,C,HW,K,ptr(D),ptr(beta),0.01,ptr(coef),None,ptr(ws),ws_bytes,ptr(packed2),st),"f")
for tile in (0,64):
    lib.clr_set_tunable(b"disc_tile",tile)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30): run(i)
    b.record(); torch.cuda.synchronize()
    print("tile",tile, a.elapsed_time(b)/30*1e3,"us per call (3 kernels)")
