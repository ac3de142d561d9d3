"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.

CPU oracle for the DeTikZify v1 image-conditioned decode path.

PARITY STATUS: **pinned by the reference's own model code for everything except the ViT internals.** The reference ships
no tests or golden vectors and its package ``__init__`` does not import here (transformers 5.5.0 vs
the pinned ~=4.52.4; timm/pymupdf/pdf2image/datasets absent), but its v1 model module does run when loaded directly:
``tests/golden/make_reference_golden.py`` executes ``detikzify/model/v1/modeling_detikzify.py`` from /root/reference
(``DetikzifyForCausalLM.forward`` / ``prepare_inputs_for_generation`` / ``generate`` with the kwargs of
infer/generate.py:218-227) on the tiny fixture weights and commits the outputs as ``tests/golden/reference_v1_tiny.pt``;
``tests/test_cpu_oracle.py`` holds this oracle to them (logits 2e-5 in fp32, greedy ids equal). What remains unpinned: the
ViT arithmetic — ``timm`` (pinned ~=1.0.11, pyproject.toml:10-12,46-48) is not installed, so both the golden script and this
oracle use HF ``SiglipVisionModel``, the published SigLIP graph timm's ``vit_so400m_patch14_siglip_384`` implements.
This oracle (i) uses the *installed* ``transformers`` 5.5.0 ``LlamaForCausalLM`` / ``SiglipVisionModel`` as the published
algorithm and (ii) restates, line by line, the reference's own glue around them:

  * concat-3 of consecutive patch tokens  .. detikzify/model/v1/modeling_detikzify.py:132-137
  * biased projector ``mm_projector``     .. detikzify/model/v1/modeling_detikzify.py:82,163
  * splice over the patch-token span + validation .. v1/modeling_detikzify.py:167-189
  * ``logits.float()``                    .. v1/modeling_detikzify.py:257
  * last-token decode with cache          .. v1/modeling_detikzify.py:285-305
  * generate() kwargs (bad words, begin-suppress, T/top-p/top-k, max_length)
                                          .. detikzify/infer/generate.py:209-227,379-387
  * SelfSim "cos" reward                  .. detikzify/evaluate/imagesim.py:91-125

The vision tower stand-in for timm ``vit_so400m_patch14_siglip_384`` is HF ``SiglipVisionModel``
(same graph; only the weight *layout* differs: timm fuses qkv and splits the MAP head's q/kv).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from transformers import (
    LlamaConfig,
    LlamaForCausalLM,
    SiglipVisionConfig,
    SiglipVisionModel,
)
from transformers.generation.logits_process import (
    LogitsProcessorList,
    NoBadWordsLogitsProcessor,
    SuppressTokensAtBeginLogitsProcessor,
    TemperatureLogitsWarper,
    TopKLogitsWarper,
    TopPLogitsWarper,
)

VPREFIX = "model.vision_model."


def _get(cfg, name, default=None):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name, default)


class Oracle:
    """fp32 (or bf16) CPU model assembled from stock HF modules + restated reference glue."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.float32,
                 attn_implementation: str = "eager"):
        vc = cfg["vision_config"]
        self.cfg = cfg
        self.dtype = dtype
        self.image_token_id = cfg["patch_token_id"]
        self.eos_token_id = cfg["eos_token_id"]
        self.concat = cfg["concat_patches"]
        self.num_patches = (vc["image_size"] // vc["patch_size"]) ** 2 // self.concat

        lcfg = LlamaConfig(
            hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
            num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
            num_key_value_heads=cfg["num_key_value_heads"], head_dim=cfg["head_dim"],
            vocab_size=cfg["vocab_size"], max_position_embeddings=cfg["max_position_embeddings"],
            rms_norm_eps=cfg["rms_norm_eps"], rope_theta=cfg["rope_theta"],
            rope_scaling=({"rope_type": "llama3", "factor": cfg["rope_factor"], "low_freq_factor": cfg["rope_low_freq_factor"],
                           "high_freq_factor": cfg["rope_high_freq_factor"],
                           "original_max_position_embeddings": cfg["rope_original_max_position"]}
                          if cfg.get("rope_type", "linear") == "llama3" else
                          {"type": "linear", "factor": cfg["rope_factor"]} if cfg["rope_factor"] != 1.0 else None),
            hidden_act="silu", attention_bias=False, mlp_bias=False, tie_word_embeddings=False,
            bos_token_id=cfg["bos_token_id"], eos_token_id=cfg["eos_token_id"],
            pad_token_id=cfg["pad_token_id"], attn_implementation=attn_implementation,
        )
        with torch.device("meta"):
            llm = LlamaForCausalLM(lcfg)
        llm = llm.to_empty(device="cpu")
        sd_llm = {k: v.to(dtype) for k, v in state_dict.items()
                  if not k.startswith(VPREFIX) and "mm_projector" not in k}
        missing, unexpected = llm.load_state_dict(sd_llm, strict=False)
        assert not unexpected, unexpected
        assert all("rotary" in m or "inv_freq" in m for m in missing), missing
        # buffers (inv_freq) were created on meta -> rebuild the rotary module on CPU
        llm.model.rotary_emb = type(llm.model.rotary_emb)(lcfg)
        self.llm = llm.to(dtype).eval()

        vcfg = SiglipVisionConfig(
            hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
            num_hidden_layers=vc["num_hidden_layers"], num_attention_heads=vc["num_attention_heads"],
            image_size=vc["image_size"], patch_size=vc["patch_size"], num_channels=vc["num_channels"],
            layer_norm_eps=vc["layer_norm_eps"], hidden_act=vc["hidden_act"],
            attn_implementation=attn_implementation,
        )
        vit = SiglipVisionModel(vcfg)
        sd_vit = {k[len(VPREFIX):]: v.to(dtype) for k, v in state_dict.items() if k.startswith(VPREFIX)}
        vit.load_state_dict(sd_vit, strict=True)
        self.vit = vit.to(dtype).eval()

        self.proj_w = state_dict["model.mm_projector.weight"].to(dtype)
        b = state_dict.get("model.mm_projector.bias")
        self.proj_b = None if b is None else b.to(dtype)

    # ---- a2/a3: ViT tokens + pooled vector (v1/modeling_detikzify.py:63-72) -----------------
    @torch.no_grad()
    def vision(self, pixel_values: torch.Tensor):
        out = self.vit(pixel_values=pixel_values.to(self.dtype))
        return out.last_hidden_state, out.pooler_output

    # ---- a4/a5: concat-3 + projector (v1/modeling_detikzify.py:132-137,163) ----------------
    @torch.no_grad()
    def image_embeds(self, pixel_values: torch.Tensor) -> torch.Tensor:
        feats, _ = self.vision(pixel_values)
        n_patch, concat = self.num_patches, self.concat
        # "in case the number of feature vectors is not divisible ... remove the first feature(s)"
        feats = feats[:, -n_patch * concat:].reshape(-1, n_patch, feats.shape[-1] * concat)
        return F.linear(feats, self.proj_w, self.proj_b)

    # ---- a6: embed + splice (v1/modeling_detikzify.py:157-189) ------------------------------
    @torch.no_grad()
    def spliced_embeds(self, input_ids: torch.Tensor, image_embeds: Optional[torch.Tensor]) -> torch.Tensor:
        inputs_embeds = self.llm.model.embed_tokens(input_ids)
        if image_embeds is None:
            return inputs_embeds
        new = []
        for i, (ids, emb) in enumerate(zip(input_ids, inputs_embeds)):
            feats = image_embeds[i]
            num = feats.shape[0]
            if (ids == self.image_token_id).sum() != num:
                raise ValueError("The number of image patch tokens should be the same as the number of image patches.")
            idx = torch.where(ids == self.image_token_id)[0]
            start = int(idx[0])
            if (idx != torch.arange(start, start + num)).any():
                raise ValueError("The image patch tokens should be consecutive.")
            new.append(torch.cat((emb[:start], feats, emb[start + num:]), dim=0))
        return torch.stack(new, dim=0)

    # ---- a7/a8: full forward -> fp32 logits for every position ------------------------------
    @torch.no_grad()
    def forward_logits(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor],
                       past_key_values=None, use_cache: bool = False):
        img = self.image_embeds(pixel_values) if pixel_values is not None else None
        embeds = self.spliced_embeds(input_ids, img)
        out = self.llm(inputs_embeds=embeds, past_key_values=past_key_values, use_cache=use_cache)
        return out.logits.float(), out.past_key_values

    @torch.no_grad()
    def decode_logits(self, token_ids: torch.Tensor, past_key_values):
        """KV-cached single-token step (v1/modeling_detikzify.py:285-305: keep last token)."""
        out = self.llm(input_ids=token_ids, past_key_values=past_key_values, use_cache=True)
        return out.logits.float(), out.past_key_values

    @torch.no_grad()
    def hidden_states(self, input_ids, pixel_values):
        img = self.image_embeds(pixel_values) if pixel_values is not None else None
        embeds = self.spliced_embeds(input_ids, img)
        out = self.llm.model(inputs_embeds=embeds, output_hidden_states=True)
        return out.hidden_states

    # ---- a10/a11: generation exactly as detikzify/infer/generate.py:218-227 drives it --------
    def logits_processors(self, prompt_len: int, temperature=None, top_p=None, top_k=None,
                          do_sample=False) -> LogitsProcessorList:
        """HF order (generation/utils.py:1109-1116,1179-1192,1213-1224): bad-words ->
        begin-suppress -> temperature -> top-k -> top-p."""
        procs = LogitsProcessorList()
        procs.append(NoBadWordsLogitsProcessor([[self.image_token_id]], eos_token_id=self.eos_token_id))
        procs.append(SuppressTokensAtBeginLogitsProcessor([self.eos_token_id], prompt_len, device="cpu"))
        if do_sample:
            if temperature is not None and temperature != 1.0:
                procs.append(TemperatureLogitsWarper(temperature))
            if top_k is not None and top_k != 0:
                procs.append(TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1))
            if top_p is not None and top_p < 1.0:
                procs.append(TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1))
        return procs

    @torch.no_grad()
    def processed_probs(self, input_ids: torch.Tensor, logits: torch.Tensor, prompt_len: int, **kw) -> torch.Tensor:
        """Post-processor probability vector (what multinomial is drawn from)."""
        scores = self.logits_processors(prompt_len, do_sample=True, **kw)(input_ids, logits.clone().float())
        return F.softmax(scores, dim=-1)

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_length: int,
                 do_sample: bool = False, temperature: float = 0.8, top_p: float = 0.95, top_k: int = 0,
                 seed: int = 0, stop_on_eos: bool = True) -> torch.Tensor:
        """Own decode loop with the processor list above (the HF ``generate`` driver is exercised
        separately in tests through ``hf_generate``). Returns [1, T] ids incl. the prompt."""
        ids = input_ids.clone()
        T0 = ids.shape[1]
        procs = self.logits_processors(T0, temperature, top_p, top_k, do_sample)
        g = torch.Generator().manual_seed(seed)
        logits, cache = self.forward_logits(ids, pixel_values, use_cache=True)
        while ids.shape[1] < max_length:
            scores = procs(ids, logits[:, -1].clone())
            if do_sample:
                nxt = torch.multinomial(F.softmax(scores, dim=-1), 1, generator=g)
            else:
                nxt = scores.argmax(dim=-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
            if stop_on_eos and int(nxt) == self.eos_token_id:
                break
            if ids.shape[1] >= max_length:
                break
            logits, cache = self.decode_logits(nxt, cache)
        return ids

    @torch.no_grad()
    def hf_generate(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_length: int,
                    **gen_kwargs) -> torch.Tensor:
        """Stock ``GenerationMixin.generate`` with the reference's kwargs (greedy or sampling)."""
        img = self.image_embeds(pixel_values) if pixel_values is not None else None
        embeds = self.spliced_embeds(input_ids, img)
        out = self.llm.generate(
            input_ids=input_ids, inputs_embeds=embeds,
            bad_words_ids=[[self.image_token_id]],
            begin_suppress_tokens=[self.eos_token_id],
            max_length=max_length, pad_token_id=self.cfg["pad_token_id"],
            **gen_kwargs)
        return out

    # ---- a13: SelfSim "cos" (detikzify/evaluate/imagesim.py:101-103,124-125) -----------------
    @torch.no_grad()
    def selfsim_cos(self, pix1: torch.Tensor, pix2: torch.Tensor) -> float:
        _, p1 = self.vision(pix1)
        _, p2 = self.vision(pix2)
        return F.cosine_similarity(p1.squeeze().double(), p2.squeeze().double(), dim=0).item()


    # ---- a13: SelfSim "emd" (detikzify/evaluate/imagesim.py:105-107,121-123): patch tokens of both images, cost 1 - cosine in
    # fp64, earth mover's distance between the UNIFORM distributions over the patches (ot.lp.emd2(M, a=[], b=[]): POT 0.9.x
    # network simplex, absent offline). Restated here as the transport LP itself (scipy HiGHS): min <P, M> s.t. P 1 = 1/n,
    # P^T 1 = 1/m, P >= 0 - independent of the assignment solver the product uses.
    @torch.no_grad()
    def selfsim_emd(self, pix1: torch.Tensor, pix2: torch.Tensor) -> float:
        import math
        import numpy as np
        from scipy.optimize import linprog
        t1, _ = self.vision(pix1)
        t2, _ = self.vision(pix2)
        a, b = t1.squeeze(0).double(), t2.squeeze(0).double()
        a = a / a.norm(dim=1, keepdim=True)
        b = b / b.norm(dim=1, keepdim=True)
        M = (1.0 - a @ b.T).numpy()
        n, m = M.shape
        A_eq = np.zeros((n + m, n * m))
        for i in range(n):
            A_eq[i, i * m:(i + 1) * m] = 1.0
        for j in range(m):
            A_eq[n + j, j::m] = 1.0
        b_eq = np.concatenate([np.full(n, 1.0 / n), np.full(m, 1.0 / m)])
        res = linprog(M.reshape(-1), A_eq=A_eq, b_eq=b_eq, bounds=(0, None), method="highs")
        assert res.status == 0, res.message
        return 2 * math.tanh(-float(res.fun)) + 1


def config_to_dict(cfg) -> dict:
    """Accept the product's DetikzifyConfig dataclass (duck-typed) or a plain dict."""
    if isinstance(cfg, dict):
        return cfg
    d = cfg.to_dict()
    return d


def synthetic_pixels(batch: int, image_size: int, seed: int = 1000) -> torch.Tensor:
    """pixel_values = 2*U[0,1)-1 (range of the (x-0.5)/0.5 normalisation), SURVEY.md §8d."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        out.append(2 * torch.rand(3, image_size, image_size, generator=g) - 1)
    return torch.stack(out)
