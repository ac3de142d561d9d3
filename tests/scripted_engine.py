"""A scripted stand-in for the CUDA engine (TEST ONLY): drives the *host* generation logic of
``DetikzifyForCausalLM.generate`` on CPU. Next-token rule: a fixed pseudo-random function of
(previous token, position) with masks applied like the real sampler; records every call."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class ScriptedEngine:
    def __init__(self, cfg, max_len: Optional[int] = None, eos_at: Optional[int] = None, newline_every: int = 5):
        self.cfg, self.device, self.max_len = cfg, torch.device("cpu"), max_len or cfg.model_max_length
        self.Vocab, self.P, self.H = cfg.vocab_size, cfg.num_patches, cfg.hidden_size
        self.calls: List[tuple] = []
        self.eos_at, self.newline_every = eos_at, newline_every
        self._slots, self._hist = set(), {}
        self._gen = None

    # -- next-token rule ---------------------------------------------------------------------
    def _next(self, prev: int, pos: int, params, suppress: bool) -> int:
        if self.eos_at is not None and pos >= self.eos_at and not suppress:
            return self.cfg.eos_token_id
        if pos % self.newline_every == 0:
            return 257  # ";\n" in the synthetic tokenizer
        tok = (prev * 31 + pos * 17 + 7) % 200 + 32
        if tok in (params.bad_token, self.cfg.eos_token_id):
            tok += 1
        return tok

    # -- engine surface used by modeling.py -----------------------------------------------------
    def seq_alloc(self):
        s = len(self._slots); self._slots.add(s); return s

    def seq_free(self, s):
        self._slots.discard(s)

    def image_embeds(self, pix):
        self.calls.append(("image_embeds", tuple(pix.shape)))
        return torch.zeros(pix.shape[0], self.P, self.H)

    def vit_encode(self, pix, want_tokens=True, want_pooled=True):
        self.calls.append(("vit_encode", tuple(pix.shape)))
        B = pix.shape[0]
        feat = pix.reshape(B, -1)[:, : self.cfg.vision_config.hidden_size].float()
        return (torch.zeros(B, self.cfg.vision_config.num_positions, self.cfg.vision_config.hidden_size) if want_tokens else None,
                feat if want_pooled else None)

    def prefill(self, slot, ids, start_pos=0, img_embeds=None, img_start=0, want_all_logits=False):
        ids = ids.tolist()
        self.calls.append(("prefill", slot, start_pos, len(ids), img_embeds is not None))
        hist = self._hist.get(slot, [])[:start_pos] + ids
        self._hist[slot] = hist
        return torch.tensor([float(hist[-1]), float(len(hist))]), None   # "logits" = (last token, length)

    @staticmethod
    def sampling(**kw):
        from types import SimpleNamespace
        return SimpleNamespace(**kw)

    def sample(self, logits, params, suppress=None, steps=None, seq_ids=None, want_probs=False):
        self.last_sampling = dict(vars(params))   # what the caller's generation kwargs became
        prev, n = int(logits[0]), int(logits[1])
        self.calls.append(("sample", bool(suppress and suppress[0])))
        return torch.tensor([self._next(prev, n, params, bool(suppress and suppress[0]))]), None

    def gen_begin(self, slots, positions, first_ids, params, seq_ids=None):
        self.calls.append(("gen_begin", positions[0], first_ids[0]))
        self._gen = dict(slot=slots[0], pos=positions[0], tok=first_ids[0], params=params, out=[])

    def gen_step(self):
        g = self._gen
        self._hist[g["slot"]] = self._hist[g["slot"]][: g["pos"]] + [g["tok"]]
        nxt = self._next(g["tok"], g["pos"] + 1, g["params"], False)
        g["pos"] += 1; g["tok"] = nxt; g["out"].append(nxt)
        self.calls.append(("gen_step",))

    def gen_wait(self, step):
        return [self._gen["out"][step]]

    def gen_end(self):
        self.calls.append(("gen_end",))
        self._gen = None

    def close(self):
        pass
