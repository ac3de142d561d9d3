from .pipeline import DetikzifyGenerator, DetikzifyPipeline, DynMinMaxNorm, NodeState, WideNode
from .tikz import TikzDocument

__all__ = ["DetikzifyGenerator", "DetikzifyPipeline", "DynMinMaxNorm", "NodeState", "WideNode", "TikzDocument"]
