"""A/B sweep of the persistent decode kernel's dev options in ONE process on one box (dev tool).

usage: decode_ab.py [model] [ctx] "k=v,k=v;k=v;..."   (each ';'-separated group is one configuration; the empty group = defaults)
Prints ms/token per configuration (three interleaved repetitions) and the max logits difference against the defaults.
"""
import subprocess
import sys

import torch

sys.path.insert(0, ".")
from detikzify_b200.model import load

name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-1.3b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
groups = sys.argv[3].split(";") if len(sys.argv) > 3 else [""]
KEYS = ["mega_variant", "mega_nslots"]
DEFAULTS = {}

model, _ = load(name, device_map=0)
eng = model.engine
slot = eng.seq_alloc()
ids = torch.randint(0, 30000, (ctx,), generator=torch.Generator().manual_seed(1)).cuda()
eng.prefill(slot, ids, 0, None, 0)
tok = torch.tensor([5], device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def apply(group):
    for k in KEYS:
        eng.set_option(k, 0)
    for k, v in DEFAULTS.items():
        eng.set_option(k, v)
    for kv in filter(None, group.split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))


print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active",
                      "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
apply("")
ref = eng.decode([slot], [ctx], tok)[0].clone()
best = {}
for rep in range(3):
    for g in groups:
        apply(g)
        for _ in range(5):
            eng.decode([slot], [ctx], tok)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(30):
            eng.decode([slot], [ctx], tok)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 30
        best[g] = min(best.get(g, 1e9), ms)
        diff = (eng.decode([slot], [ctx], tok)[0] - ref).abs().max().item()
        print(f"ctx {ctx} [{g or 'defaults'}]: ms/token {ms:.4f}  max |dlogits| vs defaults {diff:.2e}")
print("best of 3:")
for g in groups:
    print(f"  [{g or 'defaults'}]: {best[g]:.4f}")
