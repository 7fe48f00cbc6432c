import torch, time
x=torch.empty(119453696,dtype=torch.uint8).pin_memory(); d=torch.empty_like(x,device='cuda')
y=torch.empty(41944064,dtype=torch.uint8,device='cuda'); hy=torch.empty(41944064,dtype=torch.uint8).pin_memory()
for _ in range(3): d.copy_(x,non_blocking=True); hy.copy_(y,non_blocking=True)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D GB/s", 10*x.numel()/dt/1e9)
t=time.perf_counter()
for _ in range(10): hy.copy_(y,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("D2H GB/s", 10*y.numel()/dt/1e9)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(x,non_blocking=True)
    with torch.cuda.stream(s2): hy.copy_(y,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("bidir: H2D GB/s", 10*x.numel()/dt/1e9, "D2H GB/s", 10*y.numel()/dt/1e9)
