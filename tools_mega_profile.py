"""Phase-level timing breakdown of the persistent decode kernel (dev tool; run on the GPU box)."""
import ctypes as C, sys, json
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load
name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-1.3b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
model, _ = load(name, device_map=0)
eng, cfg = model.engine, model.config
slot = eng.seq_alloc()
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 30000, (ctx,), generator=g).cuda()
eng.prefill(slot, ids, 0, None, 0)
tok = torch.tensor([5], device="cuda")
for _ in range(5):
    eng.decode([slot], [ctx], tok)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    eng.decode([slot], [ctx], tok)
ev1.record(); torch.cuda.synchronize()
print("ms/token (incl tok copy kernel):", ev0.elapsed_time(ev1) / 20)
for depth in (1, 2, 3, 4, 6):
    eng.set_option('mega_depth', depth)
    for _ in range(3):
        eng.decode([slot], [ctx], tok)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        eng.decode([slot], [ctx], tok)
    ev1.record(); torch.cuda.synchronize()
    print(f'depth={depth}: ms/token {ev0.elapsed_time(ev1) / 10:.4f}')
eng.set_option('mega_depth', int(sys.argv[3]) if len(sys.argv) > 3 else 2)
for flags in (1, 2, 3, 0):
    eng.set_option("mega_flags", flags)
    for _ in range(3):
        eng.decode([slot], [ctx], tok)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        eng.decode([slot], [ctx], tok)
    ev1.record(); torch.cuda.synchronize()
    print(f"flags={flags} (1=no mma, 2=no tag waits): ms/token {ev0.elapsed_time(ev1) / 10:.4f}")
eng.set_option("mega_debug", 1)
eng.decode([slot], [ctx], tok)
L = cfg.num_hidden_layers
n = 3 * (L * 5 + 1) * 4
buf = (C.c_longlong * n)()
got = eng.lib.dtk_dbg_mega_times(eng._h, buf, n)
t = torch.tensor(list(buf), dtype=torch.float64).view(3, L * 5 + 1, 4)
names = ["qkv", "attn", "o", "gu", "down"]
for cta in range(3):
    d = t[cta]
    tot = (d[-1, 2] - d[0, 0]).item()
    print(f"CTA#{cta}: total cycles {tot:.0f}")
    for ph in range(5):
        rows = d[ph:L * 5:5]
        stage = (rows[:, 1] - rows[:, 0]).mean().item()
        items = (rows[:, 2] - rows[:, 1]).mean().item()
        print(f"  {names[ph]:5s} stage(+wait) {stage:8.0f}  items {items:8.0f}  (cycles, mean over layers)")
    lm = d[-1]
    print(f"  lm    stage {(lm[1]-lm[0]).item():8.0f}  items {(lm[2]-lm[1]).item():8.0f}")
