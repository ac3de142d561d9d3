"""Streaming-primitive microbenchmark (dev tool; run on the GPU box)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from detikzify_b200 import _lib
lib = _lib.load_library()
nbytes = 3 * 2**30
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda").random_(0, 255)
sink = torch.zeros(4, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(mode, chunk, nslots, ncw, npw, read, hint, grid=148, iters=5):
    rc = lib.dtk_dbg_stream_bench(C.c_void_p(buf.data_ptr()), nbytes, mode, chunk, nslots, ncw, npw, read, hint, grid, C.c_void_p(sink.data_ptr()), s)
    assert rc == 0, rc
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.dtk_dbg_stream_bench(C.c_void_p(buf.data_ptr()), nbytes, mode, chunk, nslots, ncw, npw, read, hint, grid, C.c_void_p(sink.data_ptr()), s)
    e1.record(); torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("LDG all warps:")
for chunk in (4096, 8192, 16384):
    for warps in (8, 12, 16):
        print(f"  chunk {chunk:6d} warps {warps:2d}: {run(1, chunk, 1, warps, 0, 0, 0):7.0f} GB/s")
print("TMA bulk ring:")
for chunk, nslots in ((4096, 32), (4096, 48), (8192, 16), (8192, 24), (16384, 8), (16384, 12), (32768, 4), (32768, 6), (65536, 3)):
    for ncw, npw in ((8, 4), (4, 4), (8, 1), (8, 8)):
        if nslots % ncw or nslots % npw:
            continue
        for read in (0, 1):
            for hint in (0, 1):
                print(f"  chunk {chunk:6d} slots {nslots:2d} ({chunk*nslots//1024:3d} KB) ncw {ncw} npw {npw} read {read} hint {hint}: {run(0, chunk, nslots, ncw, npw, read, hint):7.0f} GB/s")
