"""Small functional helpers with the reference's names (detikzify/util/functools.py:7-71)."""
from __future__ import annotations

from collections import defaultdict
from copy import copy
from functools import wraps
from typing import Any, Callable


def cache_cast(cast_func: Callable[..., Any]):
    """Memoise ``func`` on ``cast_func(*args, **kwargs)`` — lets unhashable arguments (token tensors,
    PIL images) key a cache through a user-supplied conversion (reference :7-23)."""
    def decorator(func):
        memo = {}

        @wraps(func)
        def wrapped(*args, **kwargs):
            key = cast_func(*args, **kwargs)
            if key not in memo:
                memo[key] = func(*args, **kwargs)
            return memo[key]
        return wrapped
    return decorator


def cast(cls, obj):
    clone = copy(obj)
    clone.__class__ = cls
    return clone


def listify(fn=None, wrapper=list):
    """Decorator: collect a generator's items with ``wrapper`` (reference :31-62)."""
    def deco(f):
        @wraps(f)
        def helper(*args, **kw):
            return wrapper(f(*args, **kw))
        return helper
    return deco if fn is None else deco(fn)


def batchify(fn=None):
    def to_batch(dicts):
        out = defaultdict(list)
        for d in dicts:
            for k, v in d.items():
                out[k].append(v)
        return out
    return listify(fn=fn, wrapper=to_batch)
