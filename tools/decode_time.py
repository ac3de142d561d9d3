"""ms/token of single-token decode at a given context (dev tool). DTK_B200_LIB selects an alternative build."""
import sys, subprocess
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load
name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-1.3b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
model, _ = load(name, device_map=0)
eng = model.engine
slot = eng.seq_alloc()
ids = torch.randint(0, 30000, (ctx,), generator=torch.Generator().manual_seed(1)).cuda()
eng.prefill(slot, ids, 0, None, 0)
tok = torch.tensor([5], device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for rep in range(3):
    for variant in variants:
        eng.set_option("mega_variant", variant)
        for _ in range(5):
            eng.decode([slot], [ctx], tok)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(20):
            eng.decode([slot], [ctx], tok)
        ev1.record(); torch.cuda.synchronize()
        print(f"ctx {ctx} variant {variant}: ms/token {ev0.elapsed_time(ev1) / 20:.4f}")
eng.set_option("mega_variant", 0)
lg1 = eng.decode([slot], [ctx], tok)[0].clone()
eng.set_option("decode_impl", 0)
lg0 = eng.decode([slot], [ctx], tok)[0].clone()
eng.set_option("decode_impl", 1)
print(f"persistent vs per-op logits: max abs diff {(lg1 - lg0).abs().max().item():.3e} (|logits| max {lg0.abs().max().item():.2f})")
