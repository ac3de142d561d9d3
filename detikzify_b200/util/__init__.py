from .functools import batchify, cache_cast, cast, listify
from .generation import (ExplicitAbort, StoppingCriteriaList, StreamerList, TextIteratorStreamer, TokenStreamer,
                         unwrap_processor)
from .image import DUMMY_IMAGE, expand, load, trim

__all__ = ["ExplicitAbort", "StoppingCriteriaList", "StreamerList", "TextIteratorStreamer", "TokenStreamer",
           "unwrap_processor", "cache_cast", "cast", "listify", "batchify", "DUMMY_IMAGE", "expand", "load", "trim"]
