"""End-to-end (host pointers) Viterbi throughput vs pipeline chunk size.  Each setting runs in a fresh
process because the chunk size is read once (VITERBI_B200_CHUNK_MB)."""
import json, os, subprocess, sys
code = r'''
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import viterbi_dll_b200 as vb
from viterbi_dll_b200 import dabgen
n, f = 65536, 768
syms, _ = dabgen.make_frames_torch(n, f, 3.0, seed=1, device="cuda")
h = torch.empty(syms.shape, dtype=torch.uint8, pin_memory=True); h.copy_(syms)
o = torch.empty((n, f // 8), dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
for _ in range(3): vb.lib.viterbi_deconvolve_batch(f, h.data_ptr(), n, o.data_ptr())
t0 = time.perf_counter()
for _ in range(10): vb.lib.viterbi_deconvolve_batch(f, h.data_ptr(), n, o.data_ptr())
dt = (time.perf_counter() - t0) / 10
print(round(dt * 1e3, 3), round(n * f / dt / 1e9, 2))
'''
for mb in (2048, 4096, 8192, 12288, 16384, 32768, 65536):
    env = dict(os.environ, VITERBI_B200_CHUNK_FRAMES=str(mb))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(json.dumps({"chunk_frames": mb, "ms_gbps": out.stdout.strip() or out.stderr[-200:]}), flush=True)
