import torch, time, sys
sys.path.insert(0, '/root/repo')
R,L,NY,NX=18,36,970,1042
cube=torch.rand((R,L,NY,NX),device='cuda')
out=torch.empty((L,NY,NX),device='cuda')
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: torch.sum(cube,dim=0,out=out))
print('torch sum dim0 ms',ms,'GB/s',cube.numel()*4/ms/1e6)
c2=torch.empty_like(cube)
ms=t(lambda: c2.copy_(cube))
print('copy ms',ms,'GB/s',2*cube.numel()*4/ms/1e6)
ms=t(lambda: torch.clamp_(c2,0.1,0.9))
print('inplace clamp ms',ms,'GB/s',2*cube.numel()*4/ms/1e6)
