"""SECOND, INDEPENDENT CPU RESTATEMENT — TEST INFRASTRUCTURE ONLY (pins oracle/k2_oracle.py).

The reference cannot run in this image (C# over the un-vendored Microsoft.ML.OnnxRuntime 1.22.1, ref
K2TransducerAsr.csproj:14) and holds no golden vectors, so the numpy oracle is pinned the only way that is
left: against a second implementation that shares no code with it.

* model math: the PyTorch modules the ONNX graphs are exported from (icefall `Decoder`: nn.Embedding ->
  nn.Conv1d(D, D, kernel=ctx, groups=D//4, bias=False) -> ReLU -> decoder_proj nn.Linear; `Joiner`:
  output_linear(tanh(enc + dec)); encoder_proj nn.Linear), evaluated with torch.nn.functional - where
  k2_oracle.py uses a hand-written einsum / sgemm;
* search: the textbook form - every score of a stream is sorted (value descending, flat index descending),
  hypotheses live in a dict keyed by the token tuple - where k2_oracle.py uses np.partition + lexsort on the
  top-(k+1) and an index list;
* greedy / CTC: literal per-frame Python loops over `torch.max`-free scalar folds of the reference's own
  comparison expressions (ref OfflineRecognizer.cs:150-154, :335-336).

tests/test_oracle_pin.py asserts both agree on random and crafted cases.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class TorchModel:
    def __init__(self, w: dict, blank_id: int = 0, unk_id: int = 2, context_size: int = 2, neg_id_wrap: bool = False):
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        self.emb, self.conv_w = t(w["emb"]), t(w["conv_w"])
        self.dec_w, self.dec_b = t(w["dec_proj_w"]), t(w["dec_proj_b"])
        self.out_w, self.out_b = t(w["out_w"]), t(w["out_b"])
        self.enc_w, self.enc_b = t(w.get("enc_proj_w")), t(w.get("enc_proj_b"))
        self.blank_id, self.unk_id, self.ctx, self.neg_id_wrap = blank_id, unk_id, context_size, neg_id_wrap
        self.V, self.D = self.emb.shape

    # icefall Decoder.forward (need_pad=False): embedding(clamp(y, 0)) * (y >= 0) -> permute -> conv -> permute -> relu -> proj
    def decoder(self, y) -> torch.Tensor:
        y = torch.as_tensor(np.asarray(y, dtype=np.int64)).reshape(-1, self.ctx)
        if self.neg_id_wrap:
            e = F.embedding(torch.where(y < 0, y + self.V, y), self.emb)
        else:
            e = F.embedding(y.clamp(min=0), self.emb) * (y >= 0).unsqueeze(-1).to(torch.float32)
        c = F.conv1d(e.permute(0, 2, 1), self.conv_w, bias=None, groups=self.D // 4)      # [N, D, 1]
        return F.linear(F.relu(c.permute(0, 2, 1)).squeeze(1), self.dec_w, self.dec_b)

    def joiner(self, enc, dec) -> torch.Tensor:
        enc = torch.as_tensor(np.asarray(enc, dtype=np.float32)) if not torch.is_tensor(enc) else enc
        return F.linear(torch.tanh(enc + dec), self.out_w, self.out_b)

    def encoder_proj(self, raw) -> torch.Tensor:
        return F.linear(torch.as_tensor(np.asarray(raw, dtype=np.float32)), self.enc_w, self.enc_b)


def argmax_hi_literal(row: Sequence[float]) -> int:
    """ref OfflineRecognizer.cs:150-154 as written."""
    tok = 0
    for k in range(1, len(row)):
        tok = tok if row[tok] > row[k] else k
    return tok


def greedy_single(m: TorchModel, enc: np.ndarray, max_sym_per_frame: int = 1, extra_mask: Optional[int] = None,
                  hyp: Optional[Sequence[int]] = None) -> Tuple[List[int], List[int]]:
    """ref OfflineRecognizer.cs:93-187 (single stream); with `hyp` / `extra_mask` = one stream of the online loop
    (ref OnlineRecognizer.cs:85-219). Returns (emitted tokens, timestamps)."""
    enc_t = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
    ctx = list(hyp) if hyp is not None else [-1, m.blank_id]
    toks: List[int] = []
    ts: List[int] = []
    d = m.decoder([ctx])
    t, spf = 0, 0
    while t < enc_t.shape[0] and len(toks) < 1000:
        if spf >= max_sym_per_frame:
            spf = 0
            t += 1
            continue
        y = argmax_hi_literal(m.joiner(enc_t[t:t + 1], d)[0].tolist())
        if y != m.blank_id and y != m.unk_id and y != extra_mask:
            toks.append(y)
            ts.append(t)
            ctx = (ctx + [y])[-m.ctx:]
            d = m.decoder([ctx])
            spf += 1
        else:
            spf = 0
            t += 1
    return toks, ts


def beam_search(m: TorchModel, enc: np.ndarray, beam: int, hotwords=None, hot_score: float = 0.0) -> Tuple[List[int], List[int], float]:
    """Textbook modified_beam_search of ONE stream, enc [T,J]. Returns (emitted tokens, timestamps, log-prob).
    hotwords (token-id sequences) + hot_score: contextual biasing restated WITHOUT an automaton - a hypothesis' total boost is a
    function of its token sequence (k2transducerasr_b200.hotwords.brute_force_boost plus the boost of the match in progress); an
    extension adds the difference of the totals; the match in progress is revoked at the end."""
    def boost(ys, final):
        if not hotwords:
            return 0.0
        words = [tuple(int(x) for x in w) for w in hotwords if len(w)]
        earned, hist = 0.0, ()
        for y in ys[m.ctx:]:
            hist = hist + (int(y),)
            while hist and not any(w[:len(hist)] == hist for w in words):
                hist = hist[1:]
            if hist in words:
                earned += len(hist) * hot_score
                hist = ()
        return earned + (0.0 if final else len(hist) * hot_score)
    enc_t = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
    seed = tuple([-1] * (m.ctx - 1) + [m.blank_id])
    beams: Dict[tuple, Tuple[np.float32, List[int]]] = {seed: (np.float32(0.0), [])}      # insertion-ordered
    for t in range(enc_t.shape[0]):
        keys = list(beams.keys())
        d = m.decoder([list(k[-m.ctx:]) for k in keys])
        lp = F.log_softmax(m.joiner(enc_t[t:t + 1].expand(len(keys), -1), d), dim=-1)
        lp = (lp + torch.tensor([float(beams[k][0]) for k in keys], dtype=torch.float32).unsqueeze(1)).reshape(-1)
        scored = sorted(((float(v), i) for i, v in enumerate(lp.tolist())), key=lambda p: (-p[0], -p[1]))
        new: Dict[tuple, Tuple[np.float32, List[int]]] = {}
        for v, i in scored[:beam]:
            par, tok = divmod(i, m.V)
            ys, tss = keys[par], beams[keys[par]][1]
            if tok != m.blank_id and tok != m.unk_id:
                v = float(np.float32(np.float32(v) + np.float32(boost(ys + (tok,), False) - boost(ys, False))))
                ys, tss = ys + (tok,), tss + [t]
            if ys in new:
                a, b = np.float32(new[ys][0]), np.float32(v)
                hi, lo = (a, b) if a >= b else (b, a)
                new[ys] = (np.float32(hi + np.float32(math.log1p(math.exp(float(np.float32(lo - hi)))))), new[ys][1])
            else:
                new[ys] = (np.float32(v), tss)
        beams = new
    best, best_norm, best_lp = None, None, None
    for ys, (lpv, tss) in beams.items():
        fin = np.float32(np.float32(lpv) - np.float32(boost(ys, False) - boost(ys, True)))     # revoke the match in progress
        norm = fin / np.float32(len(ys))
        if best is None or norm > best_norm:
            best, best_norm, best_lp = ys, norm, fin
    return list(best[m.ctx:]), beams[best][1], float(best_lp)


def ctc_greedy(logp: np.ndarray, blank: int = 0) -> Tuple[List[int], List[int], int]:
    """ref OfflineRecognizer.cs:328-354 for one stream, logp [T,V]: IndexOf(Max) (first maximum), collapse, trailing blanks."""
    toks: List[int] = []
    ts: List[int] = []
    prev, trailing = -1, 0
    for t in range(logp.shape[0]):
        row = logp[t].tolist()
        mx = max(row)
        y = row.index(mx)
        trailing = trailing + 1 if y == blank else 0
        if y != blank and y != prev:
            toks.append(y)
            ts.append(t)
        prev = y
    return toks, ts, trailing
