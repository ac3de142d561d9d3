"""
Canonical weight dictionary (HF parameter names) + deterministic random init.

The canonical names are those of ``transformers`` ``LlamaForCausalLM`` (decoder, ``lm_head``),
``SiglipVisionModel`` (vision tower incl. the attention-pool "map" head) plus the reference's
``model.mm_projector`` (detikzify/model/v1/modeling_detikzify.py:82). The engine packs this
dict into its device arena (see ``pack_arena`` in ``engine.py``); the test oracle loads the same
dict into stock HF modules, so both sides see bit-identical bf16-rounded parameters.

No checkpoints/tokenizers are reachable in the build environment, so ``random_init`` is the
constructor benches and tests use: matrices ~ N(0, 0.02^2)
(``initializer_range`` 0.02 in the reference, detikzify/model/configuration_detikzify.py:43),
norm gains 1 + N(0, 0.02^2), biases N(0, 0.02^2); generated in fp32 with a seeded CPU generator
and rounded once to bf16.
"""
from __future__ import annotations

from typing import Dict, Iterator, Tuple

import torch

from .configuration import DetikzifyConfig


def canonical_shapes(cfg: DetikzifyConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) with kind in {matrix, gain, bias, embed}."""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    qd = cfg.num_attention_heads * cfg.head_dim
    kd = cfg.num_key_value_heads * cfg.head_dim
    yield "model.embed_tokens.weight", (V, H), "embed"
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        yield p + "input_layernorm.weight", (H,), "gain"
        yield p + "self_attn.q_proj.weight", (qd, H), "matrix"
        yield p + "self_attn.k_proj.weight", (kd, H), "matrix"
        yield p + "self_attn.v_proj.weight", (kd, H), "matrix"
        yield p + "self_attn.o_proj.weight", (H, qd), "matrix"
        yield p + "post_attention_layernorm.weight", (H,), "gain"
        yield p + "mlp.gate_proj.weight", (I, H), "matrix"
        yield p + "mlp.up_proj.weight", (I, H), "matrix"
        yield p + "mlp.down_proj.weight", (H, I), "matrix"
    yield "model.norm.weight", (H,), "gain"
    yield "lm_head.weight", (V, H), "matrix"
    yield "model.mm_projector.weight", (H, cfg.mm_hidden_size), "matrix"
    if cfg.projector_bias:
        yield "model.mm_projector.bias", (H,), "bias"

    vc = cfg.vision_config
    D, VI, N = vc.hidden_size, vc.intermediate_size, vc.num_positions
    v = "model.vision_model.vision_model."
    yield v + "embeddings.patch_embedding.weight", (D, vc.num_channels, vc.patch_size, vc.patch_size), "matrix"
    yield v + "embeddings.patch_embedding.bias", (D,), "bias"
    yield v + "embeddings.position_embedding.weight", (N, D), "matrix"
    for l in range(vc.num_hidden_layers):
        p = v + f"encoder.layers.{l}."
        yield p + "layer_norm1.weight", (D,), "gain"
        yield p + "layer_norm1.bias", (D,), "bias"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{n}.weight", (D, D), "matrix"
            yield p + f"self_attn.{n}.bias", (D,), "bias"
        yield p + "layer_norm2.weight", (D,), "gain"
        yield p + "layer_norm2.bias", (D,), "bias"
        yield p + "mlp.fc1.weight", (VI, D), "matrix"
        yield p + "mlp.fc1.bias", (VI,), "bias"
        yield p + "mlp.fc2.weight", (D, VI), "matrix"
        yield p + "mlp.fc2.bias", (D,), "bias"
    yield v + "post_layernorm.weight", (D,), "gain"
    yield v + "post_layernorm.bias", (D,), "bias"
    h = v + "head."
    yield h + "probe", (1, 1, D), "matrix"
    yield h + "attention.in_proj_weight", (3 * D, D), "matrix"
    yield h + "attention.in_proj_bias", (3 * D,), "bias"
    yield h + "attention.out_proj.weight", (D, D), "matrix"
    yield h + "attention.out_proj.bias", (D,), "bias"
    yield h + "layernorm.weight", (D,), "gain"
    yield h + "layernorm.bias", (D,), "bias"
    yield h + "mlp.fc1.weight", (VI, D), "matrix"
    yield h + "mlp.fc1.bias", (VI,), "bias"
    yield h + "mlp.fc2.weight", (D, VI), "matrix"
    yield h + "mlp.fc2.bias", (D,), "bias"


def random_init(cfg: DetikzifyConfig, seed: int = 0, lm_head_std: float = 0.02,
                dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic parameters (CPU tensors, ``dtype``-rounded)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in canonical_shapes(cfg):
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        std = lm_head_std if name == "lm_head.weight" else 0.02
        if kind == "gain":
            t = 1.0 + 0.02 * t
        else:
            t = std * t
        out[name] = t.to(dtype)
    return out


def param_count(cfg: DetikzifyConfig) -> int:
    n = 0
    for _, shape, _ in canonical_shapes(cfg):
        k = 1
        for s in shape:
            k *= s
        n += k
    return n


# ---- timm -> canonical conversion (real v1 checkpoints ship a timm vision tower) -----------
def convert_timm_vision(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Map timm ``vit_so400m_patch14_siglip_384`` parameter names onto the canonical
    (HF SigLIP) names: timm fuses qkv into one Linear and splits the MAP head into q / kv
    projections; the math is identical."""
    v = "model.vision_model.vision_model."
    out: Dict[str, torch.Tensor] = {}
    out[v + "embeddings.patch_embedding.weight"] = sd["patch_embed.proj.weight"]
    out[v + "embeddings.patch_embedding.bias"] = sd["patch_embed.proj.bias"]
    out[v + "embeddings.position_embedding.weight"] = sd["pos_embed"].reshape(-1, sd["pos_embed"].shape[-1])
    L = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    for l in range(L):
        s, d = f"blocks.{l}.", v + f"encoder.layers.{l}."
        D = sd[s + "attn.proj.weight"].shape[0]
        for which in ("weight", "bias"):
            q, k, vv = sd[s + f"attn.qkv.{which}"].split(D, dim=0)
            out[d + f"self_attn.q_proj.{which}"] = q
            out[d + f"self_attn.k_proj.{which}"] = k
            out[d + f"self_attn.v_proj.{which}"] = vv
            out[d + f"self_attn.out_proj.{which}"] = sd[s + f"attn.proj.{which}"]
            out[d + f"layer_norm1.{which}"] = sd[s + f"norm1.{which}"]
            out[d + f"layer_norm2.{which}"] = sd[s + f"norm2.{which}"]
            out[d + f"mlp.fc1.{which}"] = sd[s + f"mlp.fc1.{which}"]
            out[d + f"mlp.fc2.{which}"] = sd[s + f"mlp.fc2.{which}"]
    for which in ("weight", "bias"):
        out[v + f"post_layernorm.{which}"] = sd[f"norm.{which}"]
        out[v + f"head.layernorm.{which}"] = sd[f"attn_pool.norm.{which}"]
        out[v + f"head.attention.out_proj.{which}"] = sd[f"attn_pool.proj.{which}"]
        out[v + f"head.mlp.fc1.{which}"] = sd[f"attn_pool.mlp.fc1.{which}"]
        out[v + f"head.mlp.fc2.{which}"] = sd[f"attn_pool.mlp.fc2.{which}"]
    out[v + "head.probe"] = sd["attn_pool.latent"]
    out[v + "head.attention.in_proj_weight"] = torch.cat(
        [sd["attn_pool.q.weight"], sd["attn_pool.kv.weight"]], dim=0)
    out[v + "head.attention.in_proj_bias"] = torch.cat(
        [sd["attn_pool.q.bias"], sd["attn_pool.kv.bias"]], dim=0)
    return out


# ---- v2 checkpoint names -> canonical (reference detikzify/model/modeling_detikzify.py:119-131: ``model.text_model`` is
# an HF LlamaModel, ``model.connector.modality_projection.proj`` the bias-free projector, ``model.vision_model`` an HF
# SiglipVisionModel; ``lm_head`` sits on the outer module) ---------------------------------------------------------------
_V2_PREFIXES = (
    ("model.text_model.", "model."),
    ("model.connector.modality_projection.proj.", "model.mm_projector."),
    ("model.vision_model.", "model.vision_model."),
)


def convert_v2_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference v2 / v2.5 parameter names -> canonical names (pure renaming, tensors are shared)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        for src, dst in _V2_PREFIXES:
            if k.startswith(src):
                out[dst + k[len(src):]] = v
                break
        else:
            out[k] = v
    return out


def to_v2_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of ``convert_v2_state_dict`` (used to load canonical test weights into the reference's v2 module)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith("model.vision_model."):
            out[k] = v
        elif k.startswith("model.mm_projector."):
            out["model.connector.modality_projection.proj." + k[len("model.mm_projector."):]] = v
        elif k.startswith("model."):
            out["model.text_model." + k[len("model."):]] = v
        else:
            out[k] = v
    return out
