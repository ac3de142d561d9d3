"""
Streamer / stopping-criterion contract of the generation boundary.

Same names, argument meaning and error behaviour as the reference's detikzify/util/generation.py
(ExplicitAbort :7-23, TokenStreamer :25-66, TextIteratorStreamer :68-79, StreamerList :81-91,
unwrap_processor :93-101), re-implemented without the HF base classes so that the hot path has no
``transformers`` dependency. Objects from the reference (or HF streamers) can be passed to
``model.generate`` interchangeably: only ``put(tensor)`` / ``end()`` / ``__call__`` are used.
"""
from __future__ import annotations

from queue import Queue
from typing import Callable, Iterable, List, Optional


class ExplicitAbort:
    """Stopping criterion flipped from another thread (reference util/generation.py:7-23)."""

    def __init__(self):
        self.should_stop = False

    def __call__(self, input_ids=None, scores=None, **kwargs) -> bool:
        return self.should_stop

    def reset(self) -> "ExplicitAbort":
        self.should_stop = False
        return self

    def abort(self) -> None:
        self.should_stop = True


class BaseStreamer:
    def put(self, value):  # pragma: no cover - interface
        raise NotImplementedError

    def end(self):  # pragma: no cover - interface
        raise NotImplementedError


class TokenStreamer(BaseStreamer):
    """Queue of raw token ids; batch size 1 only; the prompt is skipped by default."""

    def __init__(self, skip_prompt: bool = True, timeout: Optional[float] = None):
        self.skip_prompt = skip_prompt
        self.next_tokens_are_prompt = True
        self.token_queue: Queue = Queue()
        self.stop_signal = None
        self.timeout = timeout

    def put(self, value):
        shape = tuple(value.shape)
        if len(shape) > 1 and shape[0] > 1:
            raise ValueError("TokenStreamer only supports batch size 1")
        if len(shape) > 1:
            value = value[0]
        if self.skip_prompt and self.next_tokens_are_prompt:
            self.next_tokens_are_prompt = False
            return
        for token_id in value.tolist():
            self.token_queue.put(token_id, timeout=self.timeout)

    def end(self):
        self.next_tokens_are_prompt = True
        self.token_queue.put(self.stop_signal, timeout=self.timeout)

    def propagate_error(self, exc):
        self.token_queue.put(exc, timeout=self.timeout)

    def __iter__(self):
        return self

    def __next__(self):
        value = self.token_queue.get(timeout=self.timeout)
        if isinstance(value, BaseException):
            raise value
        if value is self.stop_signal:
            raise StopIteration()
        return value


class TextIteratorStreamer(BaseStreamer):
    """Decodes tokens to text incrementally and queues printable chunks (web-UI consumer)."""

    def __init__(self, tokenizer, skip_prompt: bool = False, timeout: Optional[float] = None, **decode_kwargs):
        self.tokenizer = tokenizer
        self.skip_prompt = skip_prompt
        self.decode_kwargs = decode_kwargs
        self.timeout = timeout
        self.text_queue: Queue = Queue()
        self.stop_signal = None
        self.token_cache: List[int] = []
        self.print_len = 0
        self.next_tokens_are_prompt = True

    def put(self, value):
        shape = tuple(value.shape)
        if len(shape) > 1 and shape[0] > 1:
            raise ValueError("TextIteratorStreamer only supports batch size 1")
        if len(shape) > 1:
            value = value[0]
        if self.skip_prompt and self.next_tokens_are_prompt:
            self.next_tokens_are_prompt = False
            return
        self.token_cache.extend(value.tolist())
        text = self.tokenizer.decode(self.token_cache, **self.decode_kwargs)
        if text.endswith("\n"):
            printable = text[self.print_len:]
            self.token_cache, self.print_len = [], 0
        else:
            printable = text[self.print_len: text.rfind(" ") + 1]
            self.print_len += len(printable)
        if printable:
            self.text_queue.put(printable, timeout=self.timeout)

    def end(self):
        if self.token_cache:
            text = self.tokenizer.decode(self.token_cache, **self.decode_kwargs)
            rest = text[self.print_len:]
            self.token_cache, self.print_len = [], 0
            if rest:
                self.text_queue.put(rest, timeout=self.timeout)
        self.next_tokens_are_prompt = True
        self.text_queue.put(self.stop_signal, timeout=self.timeout)

    def propagate_error(self, exc):
        self.text_queue.put(exc, timeout=self.timeout)

    def __iter__(self):
        return self

    def __next__(self):
        value = self.text_queue.get(timeout=self.timeout)
        if isinstance(value, BaseException):
            raise value
        if value is self.stop_signal:
            raise StopIteration()
        return value


class StreamerList(list, BaseStreamer):
    """Fan-out to several streamers (reference util/generation.py:81-91)."""

    def put(self, value):
        for streamer in self:
            streamer.put(value)

    def end(self):
        for streamer in self:
            streamer.end()


class StoppingCriteriaList(list):
    """Any criterion returning truthy stops generation (HF semantics for batch size 1)."""

    def __call__(self, input_ids, scores=None, **kwargs) -> bool:
        stop = False
        for crit in self:
            r = crit(input_ids, scores, **kwargs)
            try:
                stop = stop or bool(r)
            except (RuntimeError, ValueError):  # tensor with several elements
                stop = stop or bool(r.all())
        return stop


def unwrap_processor(processor):
    """Nested processors happen with the adapter processor (reference :93-101)."""
    if hasattr(processor, "processor"):
        return unwrap_processor(processor.processor)
    return processor
