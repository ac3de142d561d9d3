"""Minimal driver for `ncu --set full` captures of the non-decode kernels (GPU box):
  ncu --set full --clock-control none --import-source on -k regex:"flash_attn|gemm_tc|sample" -c 12 -o gpurun_out/r2_kernels \
      python tools/ncu_kernels.py
Runs: one ViT encode (B = 2: 128x128 tcgen05 tiles + flash attention, head_dim 72), one 32-row batched decode step of the
ds-1.3b shape (swapped-operand skinny tcgen05 tiles + split-KV attention) and the sampler (nucleus, V = 32256)."""
import sys
import torch
sys.path.insert(0, ".")
from detikzify_b200.model import load

model, _ = load("nllg/detikzify-ds-1.3b", device_map=0, max_seqs=34, max_batch=32, device_init=True)
eng, cfg = model.engine, model.config
eng.set_option("vit_graph", 0)                      # direct launches: ncu names every kernel
pix = (2 * torch.rand(2, 3, 384, 384, generator=torch.Generator().manual_seed(3)) - 1).cuda()
eng.vit_encode(pix)
ids = torch.randint(0, 30000, (300,), generator=torch.Generator().manual_seed(1)).cuda()
slots = [eng.seq_alloc() for _ in range(32)]
eng.prefill(slots[0], ids, 0, None, 0)
for s in slots[1:]:
    eng.seq_share(slots[0], s, 300)
toks = torch.full((32,), 5, device="cuda")
lg = eng.decode(slots, [300] * 32, toks)
params = eng.sampling(temperature=0.8, top_p=0.95, do_sample=True, bad_token=cfg.image_token_id, seed=3)
eng.sample(lg, params, suppress=[0] * 32, steps=list(range(32)), seq_ids=list(range(32)))
torch.cuda.synchronize()
print("done", eng.launch_count)
