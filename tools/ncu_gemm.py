"""One dense GEMM of a given shape through the kernel-level hook, for `ncu --set full` (dev tool; GPU box):
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/r2_gemm python tools/ncu_gemm.py 1000 11008 2048
Also prints the CUDA-event time of the same call (20 launches, outside any profiler range when run without ncu)."""
import ctypes as C, math, sys
import torch
sys.path.insert(0, ".")
from detikzify_b200 import _lib as L
lib = L.load_library()
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 11008, 2048)
P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
call = lambda: lib.dtk_dbg_gemm(P(A), P(W), P(None), P(None), M, N, K, 0, 0, P(None), P(out), S())
for _ in range(5):
    call()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    call()
ev1.record(); torch.cuda.synchronize()
us = ev0.elapsed_time(ev1) / 20 * 1e3
print(f"M={M} N={N} K={K} gemm_impl={lib.dtk_dbg_gemm_impl(-1)}: {us:.1f} us, {2.0 * M * N * K / us / 1e6:.1f} TF/s")
