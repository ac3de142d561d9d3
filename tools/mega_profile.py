"""Phase-level timeline of the persistent decode kernel across all CTAs (dev tool; run on the GPU box)."""
import ctypes as C, sys
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
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(label):
    for _ in range(3):
        eng.decode([slot], [ctx], tok)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        eng.decode([slot], [ctx], tok)
    ev1.record(); torch.cuda.synchronize()
    print(f"{label}: ms/token {ev0.elapsed_time(ev1) / 10:.4f}")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
timeit("default")
for flags in (1, 2, 4, 0):
    eng.set_option("mega_flags", flags)
    timeit(f"flags={flags} (1=no mma, 2=no grid barrier, 4=relaxed arrive)")
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
eng.set_option("mega_debug", 1)
eng.decode([slot], [ctx], tok)
L = cfg.num_hidden_layers
NP = L * 5 + 1
G = 148
n = G * NP * 4
buf = (C.c_longlong * n)()
got = eng.lib.dtk_dbg_mega_times(eng._h, buf, n)
G = got // (NP * 4)
t = torch.tensor(list(buf[:got]), dtype=torch.float64).view(G, NP, 4)
t0 = t[:, 0, 0].min()
names = ["qkv", "attn", "o", "gu", "down"]
print(f"kernel span: {(t[:, -1, 2].max() - t0).item() / 1e3:.1f} us  (G={G})")
for layer in (1, L // 2):
    print(f"layer {layer}: per phase over CTAs (us rel. to kernel start): start[min,max] staged[min,max] done[min,max] | mean(stage) mean(items) | n_active")
    for ph in range(5):
        d = t[:, layer * 5 + ph]
        act = (d[:, 2] - d[:, 1]) > 300  # CTAs that did work (ns)
        a = d[act] if act.any() else d
        f = lambda x: f"[{(x.min() - t0).item() / 1e3:8.2f},{(x.max() - t0).item() / 1e3:8.2f}]"
        print(f"  {names[ph]:5s} {f(a[:, 0])} {f(a[:, 1])} {f(a[:, 2])} | {(a[:, 1] - a[:, 0]).mean().item() / 1e3:6.2f} {(a[:, 2] - a[:, 1]).mean().item() / 1e3:6.2f} | {int(act.sum())}")
# slowest CTAs per phase kind (who finishes last), aggregated over layers
import collections
cnt = collections.Counter()
for layer in range(L):
    for ph in range(5):
        d = t[:, layer * 5 + ph, 2]
        cnt[int(d.argmax())] += 1
print("CTAs most often last to finish a phase:", cnt.most_common(12))

