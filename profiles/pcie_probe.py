import torch, time
n=203*1000*1000
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for _ in range(3): d.copy_(h,non_blocking=True)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h,non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D GB/s", n*10/e0.elapsed_time(e1)/1e6)
# chunked on 3 streams
ss=[torch.cuda.Stream() for _ in range(3)]
ch=32*1024*1024
e0.record()
for r in range(10):
    for i,off in enumerate(range(0,n,ch)):
        with torch.cuda.stream(ss[i%3]):
            d[off:off+ch].copy_(h[off:off+ch],non_blocking=True)
torch.cuda.synchronize(); 
t0=time.perf_counter()
for r in range(10):
    for i,off in enumerate(range(0,n,ch)):
        with torch.cuda.stream(ss[i%3]):
            d[off:off+ch].copy_(h[off:off+ch],non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("chunked 3-stream H2D GB/s", n*10/dt/1e9)
h2=torch.empty(6291456,dtype=torch.uint8,pin_memory=True); d2=torch.empty(6291456,dtype=torch.uint8,device='cuda')
t0=time.perf_counter()
for r in range(10):
    h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); print("D2H 6MB ms", (time.perf_counter()-t0)/10*1e3)
