"""Generic Monte-Carlo tree search with the interface of the MCTS package the reference vendors
(detikzify/mcts/montecarlo.py:4-85, node.py:5-68): ``MonteCarlo(root_node)`` with a pluggable
``child_finder(node, montecarlo)`` / ``node_evaluator``, ``simulate(n)`` = select by UCT until an
unexpanded node, expand it; ``Node`` with ``win_value / visits / policy_value / children / parent /
expanded / discovery_factor / is_widen_node`` and win-value back-propagation.
Host-side scalar logic (microseconds per step) — the caller of the GPU hot path, not part of it.
"""
from .montecarlo import MonteCarlo
from .node import Node

__all__ = ["MonteCarlo", "Node"]
