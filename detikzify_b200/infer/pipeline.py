"""
Sampling + MCTS driver with the reference's public surface (detikzify/infer/generate.py):

    DetikzifyPipeline(model, processor, temperature=0.8, top_p=0.95, top_k=0, compile_timeout=60,
                      metric="model"|"fast"|Metric, **gen_kwargs)
        .sample(image, text=None, preprocess=True, **gen_kwargs) -> TikzDocument          (:399-426)
        .simulate(image, text=None, preprocess=True, expansions=None, timeout=None, **kw)
              -> iterator of (score, TikzDocument)                                         (:428-464)
    DetikzifyGenerator(model, processor, image, text, metric, compile_timeout, mcts_timeout,
                       streamer, control, exploration, strict, **gen_kwargs)               (:145-353)

Search semantics kept from the reference: one expansion = one rollout from the chosen node's token
prefix (same image), a tree node per generated source line, sqrt(n) node thinning, failed-rollout memo
keyed by token prefix, error-line based pruning, min-max normalised SelfSim reward, widen nodes.
The rollout runs ``model.generate`` on a worker thread while the caller consumes a TokenStreamer —
the threading contract of SURVEY.md §8b. What is different underneath: ``model.generate`` is the B200
engine (image features cached per figure, KV prefix of the working slot reused, persistent decode
kernel), so an expansion prefills only the tree-path suffix.
"""
from __future__ import annotations

import re
from collections import deque
from dataclasses import dataclass
from functools import cached_property
from math import sqrt
from multiprocessing.pool import ThreadPool
from time import time
from types import SimpleNamespace
from typing import Any, Dict, Generator, List, Optional, Set, Tuple, Union

import torch
from PIL import Image

from ..evaluate.imagesim import ImageSim
from ..mcts import MonteCarlo, Node
from ..util import ExplicitAbort, StreamerList, TokenStreamer, cache_cast, expand, load, unwrap_processor as unwrap
from ..util.generation import StoppingCriteriaList
from .tikz import TikzDocument

Numeric = Union[int, float]


def has_adapter(model) -> bool:
    """reference detikzify/model/adapter/__init__.py:6-7 (text-conditioning adapter; not supported here)."""
    return hasattr(model, "adapter")


@dataclass(frozen=True)
class NodeState:
    token_ids: torch.Tensor
    num_lines: int = 0

    def __eq__(self, other: Any) -> bool:
        try:
            return self.token_ids.equal(other.token_ids)
        except (AttributeError, TypeError):
            return False

    def __hash__(self):
        return hash(tuple(self.token_ids.tolist()))


class WideNode(Node):
    """Tree node that always carries a "widen" twin child: selecting the twin re-expands the parent
    (progressive widening), reference :49-81."""
    state: NodeState

    def __init__(self, *args, exploration: float = 0.6, is_widen_node: bool = False, **kwargs):
        super().__init__(NodeState(*args, **kwargs))
        self.discovery_factor = exploration
        self.is_widen_node = is_widen_node
        self.update_policy_value(1.0)
        if not is_widen_node:
            self.add_child(WideNode(*args, exploration=exploration, is_widen_node=True, **kwargs))

    def add_child(self, child):
        self.expanded = self.expanded or not child.is_widen_node
        super().add_child(child)

    @property
    def depth(self) -> int:
        d, cur = 0, self
        while cur.parent is not None:
            d, cur = d + 1, cur.parent
        return d

    @property
    def token_ids(self):
        return self.state.token_ids

    @property
    def num_lines(self):
        return self.state.num_lines


class DynMinMaxNorm:
    """Scores are min-max normalised against every score seen so far, lazily (the normalisation of a
    stored value changes as new extremes arrive), reference :84-142."""

    def __init__(self, default_value: Numeric = 0):
        self.scores: Set[Numeric] = set()
        self.default_value = default_value

    def normalize(self, score: Numeric) -> "DynMinMaxNorm.MinMaxScore":
        self.scores.add(score)
        return self.MinMaxScore(score, all_scores=self.scores, default_value=self.default_value)

    __call__ = normalize

    class MinMaxScore:
        def __init__(self, *scores: Numeric, all_scores: Set[Numeric], default_value: Numeric, no_minmax_scores=()):
            self.scores = list(scores)
            self.all_scores = all_scores
            self.default_value = default_value
            self.no_minmax_scores = list(no_minmax_scores)

        @property
        def score(self) -> Numeric:
            lo, hi = min(self.all_scores), max(self.all_scores)
            if hi == lo:
                total = self.default_value
            else:
                total = sum((s - lo) / (hi - lo) for s in self.scores)
            return total + sum(self.no_minmax_scores)

        def __add__(self, other):
            new = self.__class__(*self.scores, all_scores=self.all_scores, default_value=self.default_value,
                                 no_minmax_scores=self.no_minmax_scores)
            if isinstance(other, DynMinMaxNorm.MinMaxScore):
                new.scores.extend(other.scores)
                new.no_minmax_scores.extend(other.no_minmax_scores)
            else:
                new.no_minmax_scores.append(other)
            return new

        def __mul__(self, other):
            return self.score * other

        def __truediv__(self, other):
            return self.score / other

        def __rtruediv__(self, other):
            return other / self.score

        __radd__, __rmul__ = __add__, __mul__


class DetikzifyGenerator:
    def __init__(self, model, processor, image: Optional[Image.Image], text: Optional[str] = None, metric=None,
                 compile_timeout: Optional[int] = 60, mcts_timeout: Optional[int] = None, streamer=None,
                 control: Optional[ExplicitAbort] = None, exploration: float = 0.6, strict: bool = False, **gen_kwargs):
        self.model, self.processor = model, processor
        self.metric, self.image, self.text = metric, image, text
        self.compile_timeout, self.mcts_timeout = compile_timeout, mcts_timeout
        self.streamer, self.exploration, self.strict = streamer, exploration, strict
        self.gen_kwargs = gen_kwargs
        self.solution: deque = deque(maxlen=1)
        self.failed_rollouts: Dict[NodeState, List[WideNode]] = dict()
        self.norm = DynMinMaxNorm()
        self.control = control or ExplicitAbort()
        root_ids = processor(images=self.image, text=self.text, return_tensors="pt").input_ids.to(model.device).squeeze()
        self.montecarlo = MonteCarlo(root_node=WideNode(root_ids, exploration=self.exploration))
        self.montecarlo.child_finder = self.child_finder
        self.decode = cache_cast(lambda token_ids: tuple(token_ids.tolist()))(self.decode)
        self.score = cache_cast(lambda image: image.tobytes())(self.score)

    def __call__(self, *args, **kwargs):
        return self.simulate(*args, **kwargs)

    def simulate(self, expansions: Optional[Numeric] = 1) -> Generator[Tuple[Numeric, TikzDocument], None, None]:
        """Yield every rollout (successful or not) as (score, document); reference :197-207."""
        start = time()
        while expansions is None or (expansions := expansions - 1) >= 0:
            self.montecarlo.simulate()
            yield self.solution.pop()
            if self.mcts_timeout is not None and time() - start > self.mcts_timeout:
                return

    def generate(self, input_ids: torch.Tensor, streamer=None, **gen_kwargs) -> torch.Tensor:
        """One ``model.generate`` call continuing ``input_ids`` (reference :209-227)."""
        streamers = StreamerList(filter(bool, [streamer, self.streamer]))
        numel = input_ids.numel()
        max_length = {**self.model.generation_config.to_dict(), **self.gen_kwargs, **gen_kwargs}["max_length"]
        if (numel and input_ids[-1] == unwrap(self.processor).tokenizer.eos_token_id) or numel >= max_length:
            streamers.end()
            return input_ids  # never continue past EOS / the length budget
        with torch.inference_mode():
            enc = self.processor(images=self.image, text=self.text, text_kwargs={"truncation": True}, return_tensors="pt")
            return self.model.generate(
                input_ids=input_ids.unsqueeze(0),
                bad_words_ids=[[self.model.config.image_token_id]],
                begin_suppress_tokens=[self.model.config.text_config.eos_token_id],
                pixel_values=enc.get("pixel_values"),
                streamer=streamers,
                **self.gen_kwargs,
                **gen_kwargs,
            ).squeeze()

    @cached_property
    def newlineinfo(self):
        """token id -> (number of newlines it contains, ends with newline?)  (reference :229-244)."""
        info = dict()
        for token_id in unwrap(self.processor).tokenizer.vocab.values():
            token = re.sub(r"\r\n|\r", r"\n", self.processor.decode([token_id]))
            if n := token.count("\n"):
                info[token_id] = SimpleNamespace(num_lines=n, trailing=token.endswith("\n"))
        assert info
        return info

    def rollout(self, state: NodeState) -> Generator[Tuple[torch.Tensor, int], None, None]:
        """Generate from ``state`` on a worker thread; yield (prefix ids, #lines) at every newline token."""
        input_ids, num_lines, continuation = state.token_ids, state.num_lines, False
        with ThreadPool(processes=1) as thread:
            streamer = TokenStreamer()
            result = thread.apply_async(
                func=self.generate, error_callback=streamer.propagate_error, args=[input_ids],
                kwds=dict(stopping_criteria=StoppingCriteriaList([self.control.reset()]), streamer=streamer))
            try:
                prev_ids, line = input_ids, list()
                for token in streamer:
                    line.append(token)
                    if info := self.newlineinfo.get(token):
                        num_lines += info.num_lines - continuation
                        continuation = not info.trailing
                        prev_ids = torch.cat((prev_ids, torch.tensor(line, device=prev_ids.device)))
                        line.clear()
                        yield prev_ids, num_lines
                if line:
                    yield torch.cat((prev_ids, torch.tensor(line, device=prev_ids.device))), num_lines - continuation
            except (GeneratorExit, KeyboardInterrupt):
                self.control.abort()
                raise
            else:
                if self.control.should_stop:
                    raise InterruptedError
            finally:
                result.wait()

    def decode(self, token_ids: torch.Tensor) -> TikzDocument:
        return TikzDocument(
            timeout=self.compile_timeout,
            code=self.processor.decode(token_ids=token_ids[len(self.montecarlo.root_node.token_ids):], skip_special_tokens=True))

    def score(self, image: Image.Image) -> Numeric:
        assert self.metric
        self.metric.update(img1=image, img2=self.image, text2=self.text)
        value = self.metric.compute()
        self.metric.reset()
        return value

    def sample(self):
        return self.decode(self.generate(input_ids=self.montecarlo.root_node.token_ids))

    # ---- one MCTS expansion (reference :305-343), split into its three concerns ---------------------
    def _rollout_nodes(self, start: WideNode) -> List[WideNode]:
        """Roll out from ``start`` and turn every completed source line into a candidate node; a prefix
        already known to fail short-circuits the rollout with its memoised continuation."""
        nodes: List[WideNode] = []
        stream = self.rollout(start.state)
        for ids, n_lines in stream:
            cand = WideNode(ids, n_lines, exploration=self.exploration)
            known = self.failed_rollouts.get(cand.state)
            if known is not None:
                nodes.extend(known)
                stream.close()
                break
            nodes.append(cand)
        return nodes

    def _graft(self, anchor: WideNode, nodes: List[WideNode], tikz: TikzDocument, scorable: bool) -> WideNode:
        """Attach (a thinned subset of) the rollout's nodes below ``anchor``; returns the deepest attached node."""
        keep = round(sqrt(len(nodes)))   # keep O(sqrt n) of the n lines as explicit tree nodes
        if scorable:
            for cand in nodes[:keep]:
                anchor.add_child(cand)
                anchor = cand
            return anchor
        # Failed program: only usable when the error can be located (line 0 = "unknown").
        error_line = min(tikz.errors or [0])
        if error_line:
            for idx, cand in enumerate(nodes):
                closes_line = self.newlineinfo.get(int(cand.token_ids[-1]))
                if cand.num_lines < error_line and idx < keep:
                    anchor.add_child(cand)
                    anchor = cand
                elif cand.num_lines > error_line or (cand.num_lines == error_line and closes_line):
                    self.failed_rollouts[cand.state] = nodes[idx:]
                    break
        return anchor

    def child_finder(self, node: WideNode, montecarlo: MonteCarlo):
        nodes = self._rollout_nodes(node)
        if node.is_widen_node:           # the twin stands for "sample another continuation of my parent"
            node.visits += 1
            node, nodes = self.merge(node.parent, nodes)
        tikz = self.decode((nodes or [node])[-1].token_ids)
        scorable = bool(tikz.is_rasterizable and not (self.strict and tikz.compiled_with_errors))
        node = self._graft(node, nodes, tikz, scorable)
        if self.metric:
            reward = self.score(tikz.rasterize()) if scorable else -1
        else:                            # no metric: reward from compiler diagnostics
            reward = scorable - tikz.compiled_with_errors
        node.update_win_value(self.norm(reward) if scorable and self.metric else reward)
        self.solution.append((reward, tikz))

    def merge(self, node: WideNode, nodes_to_merge: List[WideNode]) -> Tuple[WideNode, List[WideNode]]:
        """Walk down existing children that coincide with the head of the new rollout."""
        while nodes_to_merge:
            match = next((ch for ch in node.children if ch.state == nodes_to_merge[0].state), None)
            if match is None:
                break
            node, nodes_to_merge = match, nodes_to_merge[1:]
        return node, nodes_to_merge


class DetikzifyPipeline:
    def __init__(self, model, processor, temperature: float = 0.8, top_p: float = 0.95, top_k: int = 0,
                 compile_timeout: Optional[int] = 60, metric="model", **gen_kwargs):
        self.model, self.processor = model, processor
        if metric == "model":      # SelfSim
            self.metric = ImageSim.from_detikzify(model, processor, sync_on_compute=False)
        elif metric == "fast":     # compiler diagnostics only
            self.metric = None
        else:
            self.metric = metric
        self.gen_kwargs: Dict[str, Any] = dict(
            temperature=temperature, top_p=top_p, top_k=top_k,
            max_length=unwrap(processor).tokenizer.model_max_length, do_sample=True,
            compile_timeout=compile_timeout, **gen_kwargs)

    def load(self, image: Union[Image.Image, str], preprocess: bool = True):
        image = load(image)
        return expand(image, max(image.size), do_trim=True) if preprocess else image

    def check_inputs(self, image, text):
        assert text is None or has_adapter(self.model), "You need to load an adapter for textual inputs!"
        assert image or text, "Either image or text (or both) required!"

    def sample(self, image=None, text: Optional[str] = None, preprocess: bool = True, **gen_kwargs) -> TikzDocument:
        self.check_inputs(image, text)
        generator = DetikzifyGenerator(
            model=self.model, processor=self.processor,
            image=self.load(image, preprocess=preprocess) if image is not None else None, text=text,
            **self.gen_kwargs, **gen_kwargs)
        return generator.sample()

    def simulate(self, image=None, text: Optional[str] = None, preprocess: bool = True,
                 expansions: Optional[Numeric] = None, timeout: Optional[int] = None, **gen_kwargs):
        self.check_inputs(image, text)
        generator = DetikzifyGenerator(
            model=self.model, processor=self.processor, metric=self.metric, mcts_timeout=timeout or None,
            image=self.load(image, preprocess=preprocess) if image is not None else None, text=text,
            **self.gen_kwargs, **gen_kwargs)
        yield from generator.simulate(expansions or None)

    def sample_batch(self, images, preprocess: bool = True, samples_per_image: int = 1, **gen_kwargs) -> List[TikzDocument]:
        """Extension (the reference samples one figure at a time): DeTikZify several figures — and/or draw several samples
        per figure — in ONE lock-step batched decode (``model.generate_batch``): the decoder weights are streamed once per
        step for the whole batch. Returns ``len(images) * samples_per_image`` documents, image-major. Needs a model loaded
        with ``max_batch`` / ``max_seqs`` at least that large."""
        if not hasattr(self.model, "generate_batch"):
            raise TypeError("sample_batch needs a detikzify_b200 model (generate_batch)")
        images = [self.load(im, preprocess=preprocess) for im in images]
        kw = {**self.gen_kwargs, **gen_kwargs}
        timeout = kw.pop("compile_timeout", 60)
        prompts, pixels = [], []
        for im in images:
            enc = self.processor(images=im, text=None, return_tensors="pt")
            for _ in range(samples_per_image):
                prompts.append(enc.input_ids[0])
                pixels.append(enc["pixel_values"][0])
        outs = self.model.generate_batch(
            prompts, pixel_values=torch.stack(pixels), bad_words_ids=[[self.model.config.image_token_id]],
            begin_suppress_tokens=[self.model.config.text_config.eos_token_id], **kw)
        docs = []
        for ids, prompt in zip(outs, prompts):
            code = self.processor.decode(token_ids=ids[len(prompt):], skip_special_tokens=True)
            docs.append(TikzDocument(code=code, timeout=timeout))
        return docs

    def __call__(self, *args, **kwargs) -> TikzDocument:
        return self.sample(*args, **kwargs)
