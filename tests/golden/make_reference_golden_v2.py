"""
Golden vectors for the v2 wiring, produced by the REFERENCE's own v2 model code (detikzify/model/modeling_detikzify.py and
configuration_detikzify.py, read from /root/reference, never copied): ``DetikzifyForConditionalGeneration.forward`` —
SiglipVisionModel.last_hidden_state, ``DetikzifyConnector`` (concat-3 reshape + bias-free projection, :62-86),
``inputs_merger`` (masked scatter over the image-token positions, :165-179), LlamaModel with GQA and llama3 RoPE scaling,
``lm_head`` + ``.float()`` (:181-271, 329-332) — plus one KV-cached decode step through its ``prepare_inputs_for_generation``
and a greedy decode loop through its ``prepare_inputs_for_generation`` + ``forward`` with the logits processors of
infer/generate.py:218-227, on the ``tiny-v2`` fixture weights, CPU fp32.

Stubbed (recorded in DESIGN.md): the package ``__init__`` files (they import ``AutoModelForVision2Seq`` and ``datasets``,
absent here) — the two module files are loaded directly under stub parent packages; ``detikzify.model.adapter`` (the TikZero
cross-attention adapter, out of scope) is replaced by a mixin whose ``has_adapter()`` is False. transformers here is 5.5.0
(reference pins ~=4.52.4); its ``post_init`` passes a keyword the reference's 4.x-style ``tie_weights`` does not take, shimmed in
``load_reference_v2``.

Run in the build container:  python tests/golden/make_reference_golden_v2.py   -> tests/golden/reference_v2_tiny.pt
"""
import importlib.util
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/detikzify/model")

from detikzify_b200.model.configuration import preset          # noqa: E402
from detikzify_b200.model.weights import random_init, to_v2_state_dict  # noqa: E402
from oracle.hf_oracle import synthetic_pixels                   # noqa: E402


def load_reference_v2():
    for pkg in ("detikzify", "detikzify.model"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    ad = types.ModuleType("detikzify.model.adapter")

    class CrossAttentionAdapterMixin:   # out of scope (TikZero); the reference checks has_adapter() on the generate path
        def has_adapter(self):
            return False
    ad.CrossAttentionAdapterMixin = CrossAttentionAdapterMixin
    sys.modules[ad.__name__] = ad

    def load(name):
        spec = importlib.util.spec_from_file_location(f"detikzify.model.{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod
    cfgm, modm = load("configuration_detikzify"), load("modeling_detikzify")
    # transformers 5.x calls tie_weights(recompute_mapping=...) from post_init; the reference defines the 4.x signature
    # tie_weights(self). Compatibility shim for the third-party library (the reference file is untouched; with
    # tie_word_embeddings=False the method does nothing).
    orig = modm.DetikzifyForConditionalGeneration.tie_weights
    modm.DetikzifyForConditionalGeneration.tie_weights = lambda self, *a, **k: orig(self)
    return cfgm, modm


def build(name="tiny-v2", seed=0):
    cfg = preset(name)
    sd = random_init(cfg, seed=seed)
    d = cfg.to_dict()
    vc = d["vision_config"]
    cfgm, modm = load_reference_v2()
    text = dict(
        model_type="llama", hidden_size=d["hidden_size"], intermediate_size=d["intermediate_size"],
        num_hidden_layers=d["num_hidden_layers"], num_attention_heads=d["num_attention_heads"],
        num_key_value_heads=d["num_key_value_heads"], head_dim=d["head_dim"], vocab_size=d["vocab_size"],
        max_position_embeddings=d["max_position_embeddings"], rms_norm_eps=d["rms_norm_eps"], rope_theta=d["rope_theta"],
        rope_scaling={"rope_type": "llama3", "factor": d["rope_factor"], "low_freq_factor": d["rope_low_freq_factor"],
                      "high_freq_factor": d["rope_high_freq_factor"], "original_max_position_embeddings": d["rope_original_max_position"]},
        hidden_act="silu", attention_bias=False, mlp_bias=False, tie_word_embeddings=False,
        bos_token_id=d["bos_token_id"], eos_token_id=d["eos_token_id"], pad_token_id=d["pad_token_id"], attn_implementation="eager")
    vision = dict(hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"], num_hidden_layers=vc["num_hidden_layers"],
                  num_attention_heads=vc["num_attention_heads"], image_size=vc["image_size"], patch_size=vc["patch_size"],
                  num_channels=vc["num_channels"], layer_norm_eps=vc["layer_norm_eps"], hidden_act=vc["hidden_act"])
    rcfg = cfgm.DetikzifyConfig(image_token_id=d["patch_token_id"], vision_config=vision, text_config=text,
                                concat_factor=d["concat_patches"], pad_token_id=d["pad_token_id"], attn_implementation="eager")
    model = modm.DetikzifyForConditionalGeneration(rcfg).eval()
    ref_sd = {k: v.float() for k, v in to_v2_state_dict(sd).items()}
    missing, unexpected = model.load_state_dict(ref_sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    return cfg, model


def main():
    out = {}
    name = "tiny-v2"
    cfg, model = build(name)
    P = cfg.num_patches
    g = torch.Generator().manual_seed(4321)
    pre = torch.randint(0, 590, (3,), generator=g)
    post = torch.randint(0, 590, (6,), generator=g)
    ids = torch.cat([pre, torch.full((P,), cfg.patch_token_id), post]).long()
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=77)
    with torch.no_grad():
        o = model(input_ids=ids[None], pixel_values=pix, use_cache=True)
        logits = o.logits[0].float()
        nxt = int(logits[-1].argmax())
        o2 = model(input_ids=torch.tensor([[nxt]]), past_key_values=o.past_key_values, use_cache=True)
        dec = o2.logits[0, -1].float()
        img = model.model.connector(model.model.vision_model(pixel_values=pix).last_hidden_state)[0]
        prompt = torch.cat([torch.full((P,), cfg.patch_token_id), post[:3]]).long()
        # greedy generation driven the way transformers 4.x GenerationMixin drives the reference (5.x no longer passes
        # ``cache_position``, which the reference's prepare_inputs_for_generation slices input_ids with, :403-408): its own
        # prepare_inputs_for_generation + forward per step, logits processors of infer/generate.py:218-227
        from transformers import DynamicCache
        cache, gen = DynamicCache(), prompt[None].clone()
        for step in range(20):
            seen = cache.get_seq_length()
            inputs = model.prepare_inputs_for_generation(gen, past_key_values=cache, cache_position=torch.arange(seen, gen.shape[1]),
                                                         attention_mask=torch.ones_like(gen), pixel_values=pix, use_cache=True)
            lg = model(**inputs).logits[0, -1].float()
            lg[cfg.image_token_id] = -float("inf")            # bad_words_ids=[[image_token_id]]
            if step == 0:
                lg[cfg.eos_token_id] = -float("inf")          # begin_suppress_tokens=[eos]
            tok = int(lg.argmax())
            gen = torch.cat([gen, torch.tensor([[tok]])], dim=1)
            if tok == cfg.eos_token_id:
                break
        gen = gen[0]
    out[name] = {"input_ids": ids, "pixel_seed": 77, "logits": logits, "next_id": nxt, "decode_logits": dec,
                 "image_embeds": img.float(), "seed": 0, "generate_prompt": prompt, "generate_ids": gen}
    path = Path(__file__).parent / "reference_v2_tiny.pt"
    torch.save(out, path)
    print("wrote", path, {k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in out[name].items()})


if __name__ == "__main__":
    main()
