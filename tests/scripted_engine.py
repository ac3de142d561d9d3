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
        s = 0
        while s in self._slots:
            s += 1
        self._slots.add(s)
        return s

    def seq_free(self, s):
        self._slots.discard(s)

    def seq_fork(self, src, dst, length):
        self.calls.append(("seq_fork", src, dst, length))
        self._hist[dst] = list(self._hist.get(src, [])[:length])

    def seq_share(self, base, dst, length):
        self.calls.append(("seq_share", base, dst, length))
        self._hist[dst] = list(self._hist.get(base, [])[:length])

    def image_embeds(self, pix):
        self.calls.append(("image_embeds", tuple(pix.shape)))
        return torch.zeros(pix.shape[0], self.P, self.H)

    def vit_encode(self, pix, want_tokens=True, want_pooled=True):
        self.calls.append(("vit_encode", tuple(pix.shape)))
        B = pix.shape[0]
        feat = pix.reshape(B, -1)[:, : self.cfg.vision_config.hidden_size].float()
        # patch "tokens": the first hidden_size values of every position's share of the pixels (+2: pixels lie in [-1, 1], never a zero row), so that the
        # patch-token similarity modes (cos_avg, emd) see something image-dependent
        P, D = self.cfg.vision_config.num_positions, self.cfg.vision_config.hidden_size
        flat = pix.reshape(B, -1).float()
        per = flat.shape[1] // P
        tokens = flat[:, : P * per].reshape(B, P, per)[:, :, :D]
        if tokens.shape[2] < D:
            tokens = torch.nn.functional.pad(tokens, (0, D - tokens.shape[2]))
        return ((tokens + 2.0) if want_tokens else None, feat if want_pooled else None)

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
        rows = logits.view(-1, 2)                  # one ("last token", "length") row per sequence
        sup = list(suppress) if suppress else [0] * rows.shape[0]
        self.calls.append(("sample", bool(sup[0])))
        return torch.tensor([self._next(int(r[0]), int(r[1]), params, bool(sp)) for r, sp in zip(rows, sup)]), None

    def gen_begin(self, slots, positions, first_ids, params, seq_ids=None):
        self.calls.append(("gen_begin", positions[0], first_ids[0]) if len(slots) == 1 else ("gen_begin", tuple(positions), tuple(first_ids)))
        self._gen = dict(slots=list(slots), pos=list(positions), tok=list(first_ids), params=params, out=[])

    def gen_step(self):
        g = self._gen
        row = []
        for b, slot in enumerate(g["slots"]):
            pos = min(g["pos"][b], self.max_len - 1)          # the device clamps the position the same way
            self._hist[slot] = self._hist.get(slot, [])[:pos] + [g["tok"][b]]
            nxt = self._next(g["tok"][b], pos + 1, g["params"], False)
            g["pos"][b] += 1; g["tok"][b] = nxt
            row.append(nxt)
        g["out"].append(row)
        self.calls.append(("gen_step",))

    def gen_wait(self, step):
        return list(self._gen["out"][step])

    def gen_end(self):
        self.calls.append(("gen_end",))
        self._gen = None

    def close(self):
        pass
