import torch, time
x = torch.empty(2619838080//4, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device='cuda')
for n in (1, 18, 108):
    chunks = x.chunk(n); dch = d.chunk(n)
    torch.cuda.synchronize(); t=time.perf_counter()
    for a,b in zip(chunks,dch): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D', n, 'chunks', x.numel()*4/dt/1e9, 'GB/s')
y = torch.empty(145546568//4, dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize(); t=time.perf_counter(); y.copy_(d[:y.numel()], non_blocking=True); torch.cuda.synchronize(); print('D2H', y.numel()*4/(time.perf_counter()-t)/1e9)
