"""Minimal decode driver for `ncu` captures: prefill a synthetic context, then a few single-token decode steps.
Usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 4 -c 1 \
                     -o gpurun_out/r1_decode_mega python tools/ncu_decode.py nllg/detikzify-ds-1.3b 1000"""
import sys
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load

name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-1.3b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
model, _ = load(name, device_map=0)
eng = model.engine
slot = eng.seq_alloc()
ids = torch.randint(0, 30000, (ctx,), generator=torch.Generator().manual_seed(1)).cuda()
eng.prefill(slot, ids, 0, None, 0)
tok = torch.tensor([5], device="cuda")
for i in range(steps):
    eng.decode([slot], [ctx + i], tok)
torch.cuda.synchronize()
print("done", eng.launch_count)
