"""Step-by-step logits of the persistent decode kernel against the per-op decode kernels (dev tool).
usage: diag_mega.py [model] [steps] [variants]"""
import sys
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load
name = sys.argv[1] if len(sys.argv) > 1 else "tiny2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2]
model, _ = load(name, device_map=0)
eng, cfg = model.engine, model.config
slot = eng.seq_alloc()
T0 = 40
ids = torch.randint(0, cfg.vocab_size - 20, (T0,), generator=torch.Generator().manual_seed(1)).cuda()
toks = torch.randint(0, cfg.vocab_size - 20, (steps,), generator=torch.Generator().manual_seed(2)).cuda()
def run(impl, variant):
    eng.set_option("decode_impl", impl)
    eng.set_option("mega_variant", variant)
    eng.prefill(slot, ids, 0, None, 0)
    out = [eng.decode([slot], [T0 + i], toks[i:i + 1])[0].clone() for i in range(steps)]
    torch.cuda.synchronize()
    return torch.stack(out)
ref = run(0, 0)
for v in variants:
    for rep in range(3):
        got = run(1, v)
        d = (got - ref).abs().amax(dim=1)
        print(f"{name} variant {v} rep {rep}: per-step max |dlogits| vs per-op: " + " ".join(f"{x:.1e}" for x in d.tolist()), flush=True)
eng.set_option("decode_impl", 1); eng.set_option("mega_variant", 0)
