"""Driver for an ncu launch list of ONE batched decode step (dev tool; GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_batched_launches.csv \
      python tools/ncu_batched.py [model] [B] [ctx]
Direct launches (dtk_decode, not the captured graph), shared image-prefix-style KV (dtk_seq_share)."""
import sys
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load
name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-7b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 512
model, _ = load(name, device_map=0, max_seqs=B + 1, max_batch=B, device_init=True)
eng = model.engine
slots = [eng.seq_alloc() for _ in range(B)]
ids = torch.randint(0, 30000, (ctx,), generator=torch.Generator().manual_seed(1)).cuda()
eng.prefill(slots[0], ids, 0, None, 0)
for s in slots[1:]:
    eng.seq_share(slots[0], s, ctx - 16)
toks = torch.full((B,), 5, device="cuda")
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.decode(slots, [ctx] * B, toks)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
