import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import viterbi_dll_b200 as vb
from viterbi_dll_b200 import dabgen
n, f = 65536, 768
nsym, nout = 4*(f+6), f//8
syms, _ = dabgen.make_frames_torch(n, f, 3.0, seed=1, device="cuda")
h = torch.empty(syms.shape, dtype=torch.uint8, pin_memory=True); h.copy_(syms)
o = torch.empty((n, nout), dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
def run(chunk, nbuf=3):
    cp, ex, dh = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    din = [torch.empty((chunk, nsym), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    dout = [torch.empty((chunk, nout), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    free = [None]*nbuf
    def once():
        k = 0
        for lo in range(0, n, chunk):
            m = min(chunk, n-lo); b = k % nbuf
            if free[b] is not None: cp.wait_event(free[b])
            with torch.cuda.stream(cp):
                din[b][:m].copy_(h[lo:lo+m], non_blocking=True); e1 = torch.cuda.Event(); e1.record(cp)
            ex.wait_event(e1)
            with torch.cuda.stream(ex):
                vb.deconvolve_batch_device(f, din[b][:m], dout[b][:m], ex); e2 = torch.cuda.Event(); e2.record(ex)
            dh.wait_event(e2)
            with torch.cuda.stream(dh):
                o[lo:lo+m].copy_(dout[b][:m], non_blocking=True); e3 = torch.cuda.Event(); e3.record(dh)
            free[b] = e3; k += 1
        torch.cuda.synchronize()
    for _ in range(3): once()
    t0 = time.perf_counter()
    for _ in range(10): once()
    dt = (time.perf_counter()-t0)/10
    return dt
for chunk in (2048, 4096, 8192, 16384, 32768):
    dt = run(chunk)
    print(chunk, round(dt*1e3,3), "ms", round(n*f/dt/1e9,2), "Gbit/s")
# copy only
t0=time.perf_counter()
d=torch.empty_like(syms)
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("copy only", (time.perf_counter()-t0)/10*1e3)
