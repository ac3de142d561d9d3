"""Import-path compatibility with the reference (``detikzify.infer.generate``); the implementation lives in
``pipeline.py``."""
from .pipeline import *  # noqa: F401,F403
from .pipeline import DetikzifyGenerator, DetikzifyPipeline, DynMinMaxNorm, NodeState, WideNode  # noqa: F401
