"""Hot words -> the dense automaton k2b_set_context_graph takes (host side, built once per hot-word list).

The reference has no contextual biasing on its search path: `Utils/HotwordsHelper.cs:8-57` is an N-best substitution stub nobody
calls (its restatement is `text.nbest_hotwords`). What SURVEY.md section 8f rank 4 asks for is biasing INSIDE the beam search's
merge step; the semantics follow icefall's ContextGraph [EXT]: an Aho-Corasick automaton over token-id sequences in which every
matched token earns `score`, a match that breaks gives the unearned boost back, and a match still in progress at the end of the
utterance is revoked.

Definition (what the kernels, the oracle and the brute-force check below all implement):
  * trie over the hot words; node_score(n) = depth(n) * score;
  * goto(s, y): the deepest node reachable by extending a suffix of s's path with y (fail links), or the root;
  * delta(s, y) = node_score(goto(s, y)) - node_score(s)  (positive while a match deepens, negative when it falls back);
  * reaching the end of a hot word EARNS its boost: the state returns to the root and nothing is revoked later;
  * residual(s) = node_score(s): subtracted when the utterance ends in state s.
The automaton is stored densely, next / delta [S, V]: S (nodes) is tens to hundreds, V the vocabulary - the kernels look a
transition up with one load.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np


@dataclass
class ContextGraph:
    next: np.ndarray       # [S,V] int32
    delta: np.ndarray      # [S,V] float32
    residual: np.ndarray   # [S]   float32
    score: float

    @property
    def num_states(self) -> int:
        return int(self.residual.size)

    @classmethod
    def build(cls, hotwords: Sequence[Sequence[int]], vocab_size: int, score: float = 1.5) -> "ContextGraph":
        child: List[Dict[int, int]] = [{}]
        depth, is_end = [0], [False]
        for w in hotwords:
            if len(w) == 0:
                continue
            s = 0
            for y in w:
                y = int(y)
                if not 0 <= y < vocab_size:
                    raise ValueError(f"hot word token {y} outside the vocabulary")
                if y not in child[s]:
                    child[s][y] = len(child)
                    child.append({})
                    depth.append(depth[s] + 1)
                    is_end.append(False)
                s = child[s][y]
            is_end[s] = True
        S = len(child)
        # Aho-Corasick fail links, breadth first
        fail = [0] * S
        order = list(child[0].values())
        for s in order:
            fail[s] = 0
        i = 0
        while i < len(order):
            s = order[i]
            i += 1
            for y, c in child[s].items():
                f = fail[s]
                while f and y not in child[f]:
                    f = fail[f]
                fail[c] = child[f][y] if (y in child[f] and child[f][y] != c) else 0
                order.append(c)
        nxt = np.zeros((S, vocab_size), np.int32)
        dlt = np.zeros((S, vocab_size), np.float32)
        for s in [0] + order:                        # parents before children: row s can start from row fail[s]
            if s:
                nxt_raw = raw[fail[s]].copy()
            else:
                nxt_raw = np.zeros(vocab_size, np.int64)
                raw = {}
            for y, c in child[s].items():
                nxt_raw[y] = c
            raw[s] = nxt_raw
        for s in range(S):
            tgt = raw[s]
            d = (np.asarray(depth, np.float32)[tgt] - np.float32(depth[s])) * np.float32(score)
            ended = np.asarray(is_end)[tgt]
            nxt[s] = np.where(ended, 0, tgt)         # a completed hot word keeps its boost: back to the root, nothing to revoke
            dlt[s] = d
        return cls(nxt, dlt, (np.asarray(depth, np.float32) * np.float32(score)).astype(np.float32), float(score))

    def total_boost(self, tokens: Sequence[int], finalize: bool = True) -> float:
        """Sum of the deltas along `tokens` from the root (fp32, in order), minus the residual of the final state."""
        s, tot = 0, np.float32(0)
        for y in tokens:
            tot = np.float32(tot + self.delta[s, int(y)])
            s = int(self.next[s, int(y)])
        return float(tot - self.residual[s]) if finalize else float(tot)


def brute_force_boost(hotwords: Sequence[Sequence[int]], tokens: Sequence[int], score: float) -> float:
    """The definition above without an automaton: scan left to right keeping the longest suffix of the tokens seen since the last
    completed hot word that is a prefix of some hot word; a completed hot word earns len * score and clears the history."""
    words = [tuple(int(x) for x in w) for w in hotwords if len(w)]
    earned, hist = 0.0, ()
    for y in tokens:
        hist = hist + (int(y),)
        while hist and not any(w[:len(hist)] == hist for w in words):
            hist = hist[1:]
        if hist in words:
            earned += len(hist) * score
            hist = ()
    return earned
