"""Driver for an ncu launch list of ONE ViT batch (dev tool; GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/r2_vit_launches.csv python tools/ncu_vit.py [model] [B] [attn_impl] [gemm_impl]
Direct launches (vit_graph = 0), so that every kernel of the tower shows up by name."""
import sys
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load
name = sys.argv[1] if len(sys.argv) > 1 else "nllg/detikzify-ds-1.3b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
attn = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gemm = int(sys.argv[4]) if len(sys.argv) > 4 else -1
model, _ = load(name, device_map=0)
eng, cfg = model.engine, model.config
eng.set_option("vit_graph", 0)
eng.set_option("attn_impl", attn)
if gemm >= 0:
    eng.set_option("gemm_impl", gemm)
S = cfg.vision_config.image_size
pix = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(0)).mul(2).sub(1).cuda()
eng.vit_encode(pix)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
eng.vit_encode(pix)
ev1.record(); torch.cuda.synchronize()
print(f"B={B} attn_impl={attn}: {ev0.elapsed_time(ev1):.3f} ms per batch (direct launches, not under the profiler's range)")
torch.cuda.profiler.start()
eng.vit_encode(pix)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
