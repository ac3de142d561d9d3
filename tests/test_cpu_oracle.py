"""The oracle (test infrastructure) against its committed golden fixtures and against the invariants the
reference's semantics imply: KV-cached decode == full forward, concat-3 ordering, splice
validation, logits-processor order, HF generate == the oracle's own loop."""
from pathlib import Path

import pytest
import torch

from conftest import model_bundle

GOLD = torch.load(Path(__file__).parent / "golden" / "tiny_oracle.pt", weights_only=False)


def _pix(cfg, n=1, seed=1000):
    from oracle.hf_oracle import synthetic_pixels
    return synthetic_pixels(n, cfg.vision_config.image_size, seed)


@pytest.mark.parametrize("name", ["tiny", "tiny2"])
def test_oracle_matches_golden_fixtures(name):
    cfg, sd, o = model_bundle(name)
    g = GOLD[name]
    pix = _pix(cfg)
    tok, pooled = o.vision(pix)
    assert abs(tok.double().sum().item() - g["vit_tokens_sum"]) < 1e-2
    torch.testing.assert_close(tok[0, 0, :16], g["vit_tokens_row0"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(pooled[0], g["pooled"], rtol=1e-4, atol=1e-5)
    img = o.image_embeds(pix)
    torch.testing.assert_close(img[0, 0, :32], g["img_embeds_row0"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(img[0, -1, :32], g["img_embeds_last"], rtol=1e-4, atol=1e-5)
    logits, _ = o.forward_logits(g["ids"][None], pix)
    torch.testing.assert_close(logits[0, -1], g["last_logits"], rtol=1e-4, atol=1e-5)
    greedy = o.generate(g["ids"][None], pix, max_length=g["ids"].numel() + 16, stop_on_eos=False)[0]
    assert torch.equal(greedy, g["greedy"])


def test_cached_decode_equals_full_forward():
    cfg, sd, o = model_bundle("tiny")
    pix = _pix(cfg)
    ids = GOLD["tiny"]["ids"][None]
    logits, cache = o.forward_logits(ids, pix, use_cache=True)
    nxt = torch.tensor([[9]])
    step, _ = o.decode_logits(nxt, cache)
    full, _ = o.forward_logits(torch.cat([ids, nxt], 1), pix)
    assert (step[:, -1] - full[:, -1]).abs().max() < 1e-5


def test_concat3_takes_last_tokens_row_major():
    """[B, N, D] -> last 3P tokens -> [B, P, 3D] (v1/modeling_detikzify.py:132-137); tiny has N=16, P=5."""
    cfg, sd, o = model_bundle("tiny")
    pix = _pix(cfg)
    tok, _ = o.vision(pix)
    P, D = cfg.num_patches, cfg.vision_config.hidden_size
    N = tok.shape[1]
    assert N == 16 and N - 3 * P == 1   # the first patch token is dropped
    manual = torch.cat([tok[0, N - 3 * P + 3 * r: N - 3 * P + 3 * r + 3].reshape(-1) for r in range(P)]).view(P, 3 * D)
    ref = torch.nn.functional.linear(manual, o.proj_w, o.proj_b)
    assert (o.image_embeds(pix)[0] - ref).abs().max() < 1e-6


def test_splice_validation_errors():
    cfg, sd, o = model_bundle("tiny")
    pix = _pix(cfg)
    P, tok = cfg.num_patches, cfg.patch_token_id
    bad_count = torch.tensor([[tok] * (P - 1) + [5, 6]])
    with pytest.raises(ValueError, match="number of image patch tokens"):
        o.forward_logits(bad_count, pix)
    gap = torch.tensor([[tok] * (P - 1) + [5, tok]])
    with pytest.raises(ValueError, match="consecutive"):
        o.forward_logits(gap, pix)


def test_processor_order_and_nucleus():
    cfg, sd, o = model_bundle("tiny")
    V = cfg.vocab_size
    torch.manual_seed(0)
    logits = torch.randn(1, V)
    logits[0, cfg.image_token_id] = 100.0
    logits[0, cfg.eos_token_id] = 50.0
    ids = torch.zeros(1, 10, dtype=torch.long)
    p_first = o.processed_probs(ids, logits, 10, temperature=0.8, top_p=0.9, top_k=0)[0]
    assert p_first[cfg.image_token_id] == 0 and p_first[cfg.eos_token_id] == 0     # bad word + begin-suppress
    p_later = o.processed_probs(torch.zeros(1, 12, dtype=torch.long), logits, 10, temperature=0.8, top_p=0.9, top_k=0)[0]
    assert p_later[cfg.image_token_id] == 0 and p_later[cfg.eos_token_id] > 0.99   # EOS only masked at the first step
    # minimal nucleus: dropping the least likely kept token leaves < top_p mass (of the pre-warp distribution)
    base = torch.softmax(torch.where(torch.arange(V) == cfg.image_token_id, -float("inf"), logits[0].clone().index_fill(0, torch.tensor([cfg.eos_token_id]), -float("inf"))) / 0.8, -1)
    kept = p_first > 0
    assert base[kept].sum() >= 0.9 - 1e-6
    assert base[kept].sum() - base[kept].min() < 0.9 + 1e-6
    assert abs(p_first.sum() - 1) < 1e-5


def test_hf_generate_equals_oracle_loop():
    cfg, sd, o = model_bundle("tiny")
    pix = _pix(cfg)
    ids = GOLD["tiny"]["ids"][None]
    a = o.generate(ids, pix, max_length=ids.shape[1] + 12)
    b = o.hf_generate(ids, pix, max_length=ids.shape[1] + 12, do_sample=False)
    assert torch.equal(a, b)
    assert a.shape[1] <= ids.shape[1] + 12   # max_length counts the prompt (reference quirk B.4)


# ---------------------------------------------------------------- pinned by the reference's own model code
REF_GOLD = torch.load(Path(__file__).parent / "golden" / "reference_v1_tiny.pt", weights_only=False)


@pytest.mark.parametrize("name", ["tiny", "tiny2"])
def test_oracle_matches_reference_model_code(name):
    """tests/golden/reference_v1_tiny.pt was produced by the reference's OWN ``DetikzifyForCausalLM`` (detikzify/model/v1/
    modeling_detikzify.py executed from /root/reference by make_reference_golden.py; only the absent ``timm`` ViT is stood
    in for by HF SigLIP): full-prompt logits with the image span in the middle of the prompt, the concat-3 vision features,
    and one ``prepare_inputs_for_generation`` + KV-cache decode step. fp32 on both sides: agreement to summation-order noise."""
    from oracle.hf_oracle import synthetic_pixels
    cfg, sd, oracle = model_bundle(name)
    g = REF_GOLD[name]
    ids = g["input_ids"][None]
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=g["pixel_seed"])
    logits, cache = oracle.forward_logits(ids, pix, use_cache=True)
    assert logits.shape[1:] == g["logits"].shape
    assert (logits[0] - g["logits"]).abs().max().item() < 2e-5
    assert int(logits[0, -1].argmax()) == g["next_id"]
    dec, _ = oracle.decode_logits(torch.tensor([[g["next_id"]]]), cache)
    assert (dec[0, -1] - g["decode_logits"]).abs().max().item() < 2e-5
    # concat-3 order: the reference's get_vision_features output equals "last n*c tokens, c consecutive tokens per row"
    tokens, _ = oracle.vision(pix)
    n, c = cfg.num_patches, cfg.concat_patches
    feats = tokens[:, tokens.shape[1] - n * c:].reshape(-1, n, tokens.shape[-1] * c)[0]
    assert (feats - g["vision_features"]).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", ["tiny", "tiny2"])
def test_oracle_generate_equals_reference_generate(name):
    """Greedy ids of the reference's own ``DetikzifyForCausalLM.generate`` called with the kwargs of
    detikzify/infer/generate.py:218-227 (bad_words_ids, begin_suppress_tokens, max_length) == the oracle's decode loop."""
    from oracle.hf_oracle import synthetic_pixels
    cfg, sd, oracle = model_bundle(name)
    g = REF_GOLD[name]
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=g["pixel_seed"])
    out = oracle.generate(g["generate_prompt"][None], pix, max_length=g["generate_ids"].numel())
    assert out[0].tolist() == g["generate_ids"].tolist()


def test_oracle_matches_reference_model_code_at_ds13b_shape():
    """BASELINE.json configs[1] shape (detikzify-ds-1.3b: H 2048, 24 layers, RoPE theta 1e5 / linear x4, so400m ViT @384,
    729 -> 243 image tokens): last-row logits of the prompt and of one KV-cached decode step from the reference's own
    DetikzifyForCausalLM (tests/golden/make_reference_golden.py --ds13b) vs the oracle, fp32 on both sides. The GPU parity test
    at this shape (tests/test_gpu_ds13b.py) compares the CUDA path with this same oracle."""
    from oracle.hf_oracle import synthetic_pixels
    gold = torch.load(Path(__file__).parent / "golden" / "reference_v1_ds13b.pt", weights_only=False)
    cfg, sd, oracle = model_bundle("nllg/detikzify-ds-1.3b")
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=gold["pixel_seed"])
    logits, cache = oracle.forward_logits(gold["input_ids"][None], pix, use_cache=True)
    d0 = (logits[0, -1] - gold["last_logits"]).abs().max().item()
    assert d0 < 2e-4, d0
    assert int(logits[0, -1].argmax()) == gold["next_id"]
    dec, _ = oracle.decode_logits(torch.tensor([[gold["next_id"]]]), cache)
    d1 = (dec[0, -1] - gold["decode_logits"]).abs().max().item()
    assert d1 < 2e-4, d1


def test_oracle_matches_reference_v2_golden():
    """v2 wiring (GQA, llama3 RoPE, bias-free connector, masked-scatter merge): the oracle reproduces what the REFERENCE's own
    v2 module (detikzify/model/modeling_detikzify.py, run by tests/golden/make_reference_golden_v2.py) computed on the tiny-v2
    fixture — all-position logits, one cached decode step, connector output, greedy ids."""
    from pathlib import Path
    gold = torch.load(Path(__file__).parent / "golden" / "reference_v2_tiny.pt", weights_only=False)["tiny-v2"]
    from conftest import model_bundle
    from oracle.hf_oracle import synthetic_pixels
    cfg, sd, oracle = model_bundle("tiny-v2", seed=gold["seed"])
    assert cfg.num_key_value_heads < cfg.num_attention_heads and cfg.rope_type == "llama3" and not cfg.projector_bias
    pix = synthetic_pixels(1, cfg.vision_config.image_size, seed=gold["pixel_seed"])
    ids = gold["input_ids"][None]
    logits, cache = oracle.forward_logits(ids, pix, use_cache=True)
    assert (logits[0] - gold["logits"]).abs().max() < 2e-5
    assert (oracle.image_embeds(pix)[0] - gold["image_embeds"]).abs().max() < 2e-5
    dec, _ = oracle.decode_logits(torch.tensor([[gold["next_id"]]]), cache)
    assert (dec[0, -1] - gold["decode_logits"]).abs().max() < 2e-5
    out = oracle.generate(gold["generate_prompt"][None], pix, max_length=gold["generate_ids"].numel())
    assert torch.equal(out[0], gold["generate_ids"])


def test_oracle_emd_selfsim_equals_assignment_formulation():
    """oracle.selfsim_emd restates the reference's "emd" SelfSim as the transport LP between uniform marginals (what POT's
    emd2(M, [], []) solves, evaluate/imagesim.py:121-123); with equally many patches the optimum is a permutation, so the
    assignment solver the product uses must give the same value on the same (oracle) patch tokens."""
    from conftest import model_bundle
    from detikzify_b200.evaluate.imagesim import ImageSim
    from oracle.hf_oracle import synthetic_pixels
    for name in ("tiny", "tiny-v2"):
        cfg, sd, oracle = model_bundle(name)
        p1 = synthetic_pixels(1, cfg.vision_config.image_size, seed=11)
        p2 = synthetic_pixels(1, cfg.vision_config.image_size, seed=12)
        t1, _ = oracle.vision(p1)
        t2, _ = oracle.vision(p2)
        assert ImageSim._emd_similarity(t1.squeeze(0), t2.squeeze(0)) == pytest.approx(oracle.selfsim_emd(p1, p2), abs=1e-9)
        assert oracle.selfsim_emd(p1, p1) == pytest.approx(1.0, abs=1e-9)
