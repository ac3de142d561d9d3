"""
Golden vectors produced by the REFERENCE's own model code (detikzify/model/v1/modeling_detikzify.py, read from
/root/reference, never copied): ``DetikzifyForCausalLM.forward`` — get_vision_features (concat-3 slice + reshape), mm_projector,
the patch-token splice loop, LlamaModel.forward, lm_head, ``logits.float()`` — and its ``prepare_inputs_for_generation`` +
KV-cache decode step, run on the tiny fixture weights on CPU in fp32.

What is stubbed, and why (recorded in DESIGN.md §6):
  * ``timm`` is not installed here. ``timm.create_model`` is replaced by a wrapper that exposes the four members the reference
    touches (``get_intermediate_layers(x, n=[layer], norm=True)``, ``forward_features``, ``forward_head``, ``patch_embed``)
    on top of HF ``SiglipVisionModel`` — the same SigLIP graph timm's ``vit_so400m_patch14_siglip_384`` implements. The ViT
    arithmetic itself is therefore pinned by HF's implementation, not timm's.
  * The package ``__init__`` files import things absent from transformers 5.x (``AutoModelForVision2Seq``) and ``datasets``;
    the three v1 module files are loaded directly under stub parent packages instead. ``processing_detikzify`` (host image
    preprocessing, needs timm.data) is stubbed; it is not part of the arithmetic.

  * ``transformers`` here is 5.5.0, the reference pins ~=4.52.4. One behavioural difference reaches this path: in 4.x an
    EMPTY ``DynamicCache`` is falsy (``__len__`` = number of populated layers), and the reference's
    ``prepare_inputs_for_generation`` relies on it (``if past_key_values: input_ids = input_ids[:, -1:]``, :288-289); 5.x
    caches are always truthy, which would make ``generate()`` drop the prompt. The script restores the 4.x ``__len__`` on
    ``DynamicCache`` while it runs (a compatibility shim for the third-party library, the reference code is untouched).

Run (in the build container, where /root/reference exists):  python tests/golden/make_reference_golden.py [--ds13b]
Writes tests/golden/reference_v1_tiny.pt (inputs + reference outputs; the test regenerates the weights from the seed);
with --ds13b, tests/golden/reference_v1_ds13b.pt at the real detikzify-ds-1.3b shape (last-row logits only).
"""
import importlib.util
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/detikzify/model/v1")

from detikzify_b200.model.configuration import preset          # noqa: E402
from detikzify_b200.model.weights import random_init           # noqa: E402
from oracle.hf_oracle import VPREFIX, synthetic_pixels          # noqa: E402


def load_reference_v1():
    from transformers import SiglipVisionConfig, SiglipVisionModel

    state = {}

    class TimmLikeSiglip(torch.nn.Module):
        """The members of a timm VisionTransformer the reference uses, over HF SiglipVisionModel."""
        def __init__(self, vcfg):
            super().__init__()
            self.hf = SiglipVisionModel(vcfg)
            self.pretrained_cfg = {"architecture": "vit_so400m_patch14_siglip_384"}
            self.embed_dim = vcfg.hidden_size
            self.blocks = list(self.hf.vision_model.encoder.layers)
            pe = self.hf.vision_model.embeddings.patch_embedding
            self.patch_embed = types.SimpleNamespace(proj=pe, num_patches=(vcfg.image_size // vcfg.patch_size) ** 2)

        def forward_features(self, x):
            return self.hf(pixel_values=x).last_hidden_state

        def forward_head(self, feats):
            return self.hf.vision_model.head(feats)

        def get_intermediate_layers(self, x, n=None, norm=True, **kw):
            # last block + final norm (feature_layer = -1 in every released checkpoint)
            assert norm and list(n) == [len(self.blocks) - 1], (n, norm)
            return [self.forward_features(x)]

    timm = types.ModuleType("timm")
    timm.create_model = lambda name, **kw: TimmLikeSiglip(state["vcfg"])
    sys.modules["timm"] = timm
    for pkg in ("detikzify", "detikzify.model", "detikzify.model.v1"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    proc = types.ModuleType("detikzify.model.v1.processing_detikzify")
    proc.DetikzifyImageProcessor = type("DetikzifyImageProcessor", (), {})
    sys.modules[proc.__name__] = proc

    def load(name):
        spec = importlib.util.spec_from_file_location(f"detikzify.model.v1.{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod
    cfgm = load("configuration_detikzify")
    modm = load("modeling_detikzify")
    return cfgm, modm, state, SiglipVisionConfig


def build(name="tiny", seed=0):
    cfg = preset(name)
    sd = random_init(cfg, seed=seed)
    d = cfg.to_dict()
    vc = d["vision_config"]
    cfgm, modm, state, SiglipVisionConfig = load_reference_v1()
    state["vcfg"] = SiglipVisionConfig(
        hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"], num_hidden_layers=vc["num_hidden_layers"],
        num_attention_heads=vc["num_attention_heads"], image_size=vc["image_size"], patch_size=vc["patch_size"],
        num_channels=vc["num_channels"], layer_norm_eps=vc["layer_norm_eps"], hidden_act=vc["hidden_act"],
        attn_implementation="eager")
    n_tokens = (vc["image_size"] // vc["patch_size"]) ** 2
    rcfg = cfgm.DetikzifyConfig(
        hidden_size=d["hidden_size"], intermediate_size=d["intermediate_size"], num_hidden_layers=d["num_hidden_layers"],
        num_attention_heads=d["num_attention_heads"], num_key_value_heads=d["num_key_value_heads"], head_dim=d["head_dim"],
        vocab_size=d["vocab_size"], max_position_embeddings=d["max_position_embeddings"], rms_norm_eps=d["rms_norm_eps"],
        rope_theta=d["rope_theta"],
        rope_scaling={"type": "linear", "factor": d["rope_factor"]} if d["rope_factor"] != 1.0 else None,
        hidden_act="silu", attention_bias=False, mlp_bias=False, tie_word_embeddings=False,
        bos_token_id=d["bos_token_id"], eos_token_id=d["eos_token_id"], pad_token_id=d["pad_token_id"],
        attn_implementation="eager",
        # the fields initialize_vision_modules() writes (v1/modeling_detikzify.py:100-107)
        use_mm_proj=True, vision_tower="vit_so400m_patch14_siglip_384.webli", mm_hidden_size=vc["hidden_size"] * d["concat_patches"],
        patch_token_id=d["patch_token_id"], concat_patches=d["concat_patches"], feature_layer=vc["num_hidden_layers"] - 1,
        num_patches=n_tokens // d["concat_patches"])
    model = modm.DetikzifyForCausalLM(rcfg).eval()
    sd_llm = {k: v.float() for k, v in sd.items() if not k.startswith(VPREFIX)}
    missing, unexpected = model.load_state_dict(sd_llm, strict=False)
    assert not unexpected, unexpected
    assert all("vision_model" in m or "rotary" in m for m in missing), missing
    vit = model.model.vision_model.model[0]
    vit.hf.load_state_dict({k[len(VPREFIX):]: v.float() for k, v in sd.items() if k.startswith(VPREFIX)}, strict=True)
    vit.eval()
    return cfg, model.float()


def restore_v4_cache_truthiness():
    from transformers import DynamicCache
    # 5.x: generate() builds DynamicCache(config=...), whose layer list is pre-sized -> truthy although it holds no tokens
    DynamicCache.__len__ = lambda self: 0 if self.get_seq_length() == 0 else len(self.layers)
    assert not DynamicCache()


@torch.no_grad()
def main():
    restore_v4_cache_truthiness()
    torch.manual_seed(0)
    out = {}
    for name in ("tiny", "tiny2"):
        cfg, model = build(name)
        P = cfg.num_patches
        g = torch.Generator().manual_seed(4242)
        n_text = 9
        text = torch.randint(0, min(cfg.vocab_size, cfg.patch_token_id), (n_text,), generator=g)
        # image span in the MIDDLE of the prompt exercises the splice offsets (reference prompts put it first)
        ids = torch.cat([text[:3], torch.full((P,), cfg.patch_token_id), text[3:]]).long()[None]
        pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=77)
        res = model(input_ids=ids, pixel_values=pix, use_cache=True, return_dict=True)
        logits = res.logits[0].clone()
        # one KV-cached decode step exactly as generate() drives it (prepare_inputs_for_generation, :285-305)
        nxt = logits[-1].argmax()[None, None]
        ids2 = torch.cat([ids, nxt], dim=1)
        inputs = model.prepare_inputs_for_generation(ids2, past_key_values=res.past_key_values, use_cache=True, pixel_values=pix)
        inputs = {k: v for k, v in inputs.items() if v is not None}
        res2 = model(**inputs, return_dict=True)
        feats = model.model.get_vision_features(pix)[0].clone()
        # the reference's generation call (detikzify/infer/generate.py:218-227) on its own model class, greedy
        prompt = torch.cat([torch.full((P,), cfg.patch_token_id), text[:4]]).long()[None]
        gen = model.generate(input_ids=prompt, pixel_values=pix, bad_words_ids=[[cfg.patch_token_id]],
                             begin_suppress_tokens=[cfg.eos_token_id], max_length=prompt.shape[1] + 24, do_sample=False,
                             pad_token_id=cfg.pad_token_id)
        out[name] = {"input_ids": ids[0], "pixel_seed": 77, "logits": logits, "next_id": int(nxt), "decode_logits": res2.logits[0, -1].clone(),
                     "vision_features": feats, "seed": 0, "generate_prompt": prompt[0], "generate_ids": gen[0].clone()}
        print(name, "logits", tuple(logits.shape), "max|logit|", float(logits.abs().max()), "decode ok")
    torch.save(out, Path(__file__).with_name("reference_v1_tiny.pt"))
    print("wrote", Path(__file__).with_name("reference_v1_tiny.pt"))


@torch.no_grad()
def main_ds13b():
    """BASELINE.json configs[1] shape: the reference model code at the real detikzify-ds-1.3b dimensions (random-init weights
    from the seed). Only the last-row logits of the prompt and of one cached decode step are kept (2 x 32256 floats)."""
    restore_v4_cache_truthiness()
    name = "nllg/detikzify-ds-1.3b"
    cfg, model = build(name)
    P = cfg.num_patches
    g = torch.Generator().manual_seed(1313)
    ids = torch.cat([torch.full((P,), cfg.patch_token_id), torch.randint(0, 32000, (5,), generator=g)]).long()[None]
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=13)
    res = model(input_ids=ids, pixel_values=pix, use_cache=True, return_dict=True)
    last = res.logits[0, -1].clone()
    nxt = last.argmax()[None, None]
    inputs = model.prepare_inputs_for_generation(torch.cat([ids, nxt], dim=1), past_key_values=res.past_key_values, use_cache=True,
                                                 pixel_values=pix)
    res2 = model(**{k: v for k, v in inputs.items() if v is not None}, return_dict=True)
    out = {"input_ids": ids[0], "pixel_seed": 13, "seed": 0, "last_logits": last, "next_id": int(nxt),
           "decode_logits": res2.logits[0, -1].clone()}
    torch.save(out, Path(__file__).with_name("reference_v1_ds13b.pt"))
    print("ds-1.3b: max|logit|", float(last.abs().max()), "next", int(nxt), "-> reference_v1_ds13b.pt")


if __name__ == "__main__":
    if "--ds13b" in sys.argv:
        main_ds13b()
    else:
        main()
