from __future__ import annotations

import math
import random
from typing import Any, List, Optional


class Node:
    """Search-tree node. Score of a child = exploitation (mean win value, zero for widen nodes)
    + exploration (discovery_factor * prior * sqrt(ln(parent visits) / visits)), cf. reference
    node.py:51-68."""

    def __init__(self, state: Any):
        self.state = state
        self.win_value: float = 0
        self.policy_value: Optional[float] = None
        self.visits: int = 0
        self.parent: Optional["Node"] = None
        self.children: List["Node"] = []
        self.expanded: bool = False
        self.player_number = None
        self.discovery_factor: float = 0.35
        self.is_widen_node: bool = False
        self.score: float = 0.0

    # -- statistics ----------------------------------------------------------------------------
    def update_win_value(self, value) -> None:
        node: Optional[Node] = self
        while node is not None:          # back-propagate to the root
            node.win_value += value
            node.visits += 1
            node = node.parent

    def update_policy_value(self, value) -> None:
        self.policy_value = value

    # -- structure -----------------------------------------------------------------------------
    def add_child(self, child: "Node") -> None:
        child.parent = self
        self.children.append(child)

    def add_children(self, children) -> None:
        for child in children:
            self.add_child(child)

    # -- selection -----------------------------------------------------------------------------
    def get_score(self, root_node: "Node") -> float:
        visits = self.visits or 1
        explore = self.discovery_factor * (self.policy_value or 1) * math.sqrt(math.log(self.parent.visits) / visits)
        if self.is_widen_node:
            exploit = 0
        else:
            sign = 1 if self.parent.player_number == root_node.player_number else -1
            exploit = sign * self.win_value / visits
        self.score = exploit + explore
        return self.score

    def get_preferred_child(self, root_node: "Node") -> "Node":
        best, best_score = [], -math.inf
        for child in self.children:
            s = child.get_score(root_node)
            if s > best_score:
                best, best_score = [child], s
            elif s == best_score:
                best.append(child)
        return random.choice(best)

    def is_scorable(self) -> bool:
        return bool(self.visits) or self.policy_value is not None
