"""Generates tests/golden/tiny_oracle.pt — small fixtures of the CPU oracle (fp32, transformers 5.5.0 at
generation time) on the seeded `tiny` / `tiny2` configurations. The reference ships no golden vectors
 and its package is not importable offline, so these pin OUR restatement against drift of
the installed HF modules and give the `-m gpu` tests committed expectations that do not need the oracle's
weights path. Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from detikzify_b200.model.configuration import preset  # noqa: E402
from detikzify_b200.model.weights import random_init  # noqa: E402
from oracle.hf_oracle import Oracle, synthetic_pixels  # noqa: E402


def main():
    out = {"transformers": __import__("transformers").__version__, "torch": torch.__version__}
    for name in ("tiny", "tiny2"):
        cfg = preset(name)
        sd = random_init(cfg, seed=0)
        o = Oracle(cfg.to_dict(), sd)
        pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=1000)
        g = torch.Generator().manual_seed(2000)
        text = torch.randint(0, min(cfg.vocab_size, cfg.patch_token_id), (7,), generator=g)
        ids = torch.cat([torch.full((cfg.num_patches,), cfg.patch_token_id), text]).long()
        tok, pooled = o.vision(pix)
        img = o.image_embeds(pix)
        logits, _ = o.forward_logits(ids[None], pix)
        greedy = o.generate(ids[None], pix, max_length=ids.numel() + 16, stop_on_eos=False)[0]
        probs = o.processed_probs(ids[None], logits[:, -1], ids.numel(), temperature=0.8, top_p=0.95, top_k=0)
        out[name] = dict(ids=ids, vit_tokens_sum=tok.double().sum().item(), vit_tokens_row0=tok[0, 0, :16].clone(),
                         pooled=pooled[0].clone(), img_embeds_row0=img[0, 0, :32].clone(), img_embeds_last=img[0, -1, :32].clone(),
                         last_logits=logits[0, -1].clone(), logits_pos0=logits[0, 0, :64].clone(), greedy=greedy,
                         nucleus_size=int((probs > 0).sum()), nucleus_probs_top=probs[0].topk(8).values.clone())
    torch.save(out, Path(__file__).with_name("tiny_oracle.pt"))
    print("wrote", Path(__file__).with_name("tiny_oracle.pt"))


if __name__ == "__main__":
    main()
