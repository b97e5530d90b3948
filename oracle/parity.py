"""PARITY CHECKER — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).

Compares what the CUDA library returned with what oracle/k2_oracle.py returned on the same inputs and produces the
numbers BASELINE.json's north_star asks for:

  * frames_identical_pct  - share of (stream, frame) cells whose emitted symbol (or "nothing") is identical;
  * near_tie_frames       - the FIRST divergent frame of every differing stream when the oracle's own decision margin
                            AT THAT FRAME is below 1e-4 (listed separately, as the north star says);
  * resolution_ties       - first divergent frames whose margin is at most 4 float32 spacings of the hypothesis scores being
                            compared (|log-prob| of a few hundred after 200+ frames has a spacing of 3e-5 .. 6e-5: a margin of
                            1.2e-4 is TWO representable steps - below what fp32 accumulation in any order can decide), listed
                            separately from the near ties;
  * unexplained           - first divergent frames that are neither: any entry is a parity failure;
  * max_score_err         - largest |log-prob difference| over the streams whose output is identical.

For modified_beam_search the divergence is located on the BEAM, not on the final output: the library's back-pointer
history (k2b_debug_backpointers: per frame and slot, parent slot + appended token) is compared frame by frame with the
oracle's `history`, so a difference is attributed to the frame where the two beams first differ - a stream cannot hide
behind a near tie it met somewhere else. The same comparison verifies the 64-bit hash dedupe against real token
sequences: the oracle merges on the token tuple, so a hash collision (wrong merge) or a missed merge shows up as a beam
difference at a frame without a near tie.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

NEAR_TIE = 1e-4      # BASELINE.json north_star: top-2 gap under 1e-4
SCORE_TOL = 1e-3     # hypothesis log-probs: 1e-3 absolute in fp32
RESOLUTION_ULPS = 4  # a margin of at most this many float32 spacings of the compared scores is undecidable in fp32


@dataclass
class ParityReport:
    streams: int = 0
    frames: int = 0
    streams_identical: int = 0
    frames_identical: int = 0
    beam_frames_compared: int = 0
    beam_frames_identical: int = 0
    near_tie_frames: List[tuple] = field(default_factory=list)   # (stream, frame, margin)
    resolution_ties: List[tuple] = field(default_factory=list)   # (stream, frame, margin): margin <= 4 ulp of the scores
    unexplained: List[tuple] = field(default_factory=list)       # (stream, frame, margin)
    output_differs: List[int] = field(default_factory=list)      # streams whose final tokens / timestamps differ
    cascade_streams: List[int] = field(default_factory=list)     # coupled batches: streams that follow another stream's tie
    max_score_err: float = 0.0

    @property
    def frames_identical_pct(self) -> float:
        return 100.0 * self.frames_identical / self.frames if self.frames else 100.0

    def as_dict(self, limit: int = 16) -> dict:
        d = {"streams": self.streams, "frames": self.frames, "streams_identical": self.streams_identical,
             "frames_identical_pct": round(self.frames_identical_pct, 4),
             "near_tie_frames": len(self.near_tie_frames),
             "near_tie_list": [[int(s), int(t), float(f"{g:.3g}")] for s, t, g in self.near_tie_frames[:limit]],
             "fp32_resolution_ties": len(self.resolution_ties),
             "fp32_resolution_list": [[int(s), int(t), float(f"{g:.3g}")] for s, t, g in self.resolution_ties[:limit]],
             "unexplained_frames": len(self.unexplained),
             "max_score_err": float(f"{self.max_score_err:.3g}")}
        if self.beam_frames_compared:
            # frames up to each stream's first beam difference (a slot swap at a tie counts every later frame as different)
            d["beam_prefix_identical_pct"] = round(100.0 * self.beam_frames_identical / self.beam_frames_compared, 4)
        return d

    def assert_ok(self, what: str = "", min_frames_pct: float = 99.9, score_tol: float = SCORE_TOL) -> None:
        assert not self.unexplained, (f"{what}: {len(self.unexplained)} stream(s) diverge from the oracle at a frame that is "
                                      f"no near tie (stream, frame, margin): {self.unexplained[:8]}")
        assert self.frames_identical_pct >= min_frames_pct, (
            f"{what}: only {self.frames_identical_pct:.3f} % of the frames are identical (near ties: {self.near_tie_frames[:8]})")
        assert self.max_score_err <= score_tol, f"{what}: hypothesis log-prob off by {self.max_score_err:.3g}"


def emit_map(tokens: Sequence[int], ts: Sequence[int], T: int, t0: int = 0) -> np.ndarray:
    """Per frame the symbols emitted on it, packed into one integer (at most a few symbols per frame)."""
    m = np.full(T, -1, np.int64)
    for tok, t in zip(tokens, ts):
        i = int(t) - t0
        assert 0 <= i < T, f"timestamp {t} outside [{t0}, {t0 + T})"
        m[i] = int(tok) if m[i] < 0 else m[i] * 1000003 + int(tok) + 1
    return m


def beam_records(bp_row: np.ndarray):
    """[K] int32 back-pointer entries of one (stream, frame) -> list of (parent slot, token or -1)."""
    return [((int(e) >> 28) & 0xF, (int(e) & 0x0FFFFFFF) - 1) for e in bp_row]


def first_beam_divergence(bp: np.ndarray, history: Sequence[Sequence[tuple]]) -> int:
    """bp [T,K] of one stream vs the oracle's per-frame records; returns the first differing frame or -1."""
    T, K = bp.shape
    for t in range(min(T, len(history))):
        got = beam_records(bp[t])
        want = list(history[t])
        if got[:len(want)] != [(int(p), int(k)) for p, k in want] or any(g != (0, -1) for g in got[len(want):]):
            return t
    return -1


def compare(got_tokens: Sequence[Sequence[int]], got_ts: Sequence[Sequence[int]], want, T: int,
            got_score: Optional[Sequence[float]] = None, bp: Optional[np.ndarray] = None, coupled: bool = False,
            t0: int = 0, near_tie: float = NEAR_TIE) -> ParityReport:
    """want: list of k2_oracle.StreamResult (with frame_gap; history when bp is given). T frames per stream (streams
    decoded over fewer frames - ragged batches - simply have no symbols beyond their length). bp: [B,T,K] int32.
    coupled: the streams influence each other (the reference's batch greedy, Q6): a stream that diverges at or after the
    batch's first, near-tie, divergence is listed as a cascade instead of being judged on its own margin."""
    rep = ParityReport(streams=len(want), frames=len(want) * T)
    assert len(got_tokens) == len(want) and len(got_ts) == len(want)
    first = []
    for b, r in enumerate(want):
        n = len(r.appended)
        want_ts = list(r.timestamps)[len(r.timestamps) - n:] if n else []
        gm = emit_map(got_tokens[b], got_ts[b], T, t0)
        wm = emit_map(r.appended, want_ts, T, t0)
        same = gm == wm
        rep.frames_identical += int(same.sum())
        out_same = bool(same.all()) and list(got_tokens[b]) == list(r.appended)
        t_div, margin = -1, float("inf")
        if not out_same:
            rep.output_differs.append(b)
        if bp is not None and r.history is not None:
            rep.beam_frames_compared += T
            t_div = first_beam_divergence(np.asarray(bp[b]), r.history)
            rep.beam_frames_identical += T if t_div < 0 else t_div
            if t_div >= 0:
                margin = r.frame_gap[t_div]
            elif not out_same:                         # identical beams, different final pick
                t_div, margin = T - 1, r.final_gap
        elif not out_same:
            t_div = int(np.argmin(same)) if not same.all() else T - 1
            if r.history is not None:      # beam search judged on its output only: any decision up to the end may be the cause
                margin = min(min(r.frame_gap, default=float("inf")), r.final_gap)
            else:
                margin = r.frame_gap[t_div] if t_div < len(r.frame_gap) else r.min_gap
        if out_same:
            rep.streams_identical += 1
            if got_score is not None:
                rep.max_score_err = max(rep.max_score_err, abs(float(got_score[b]) - float(r.score)))
        if t_div >= 0:
            first.append((b, t_div, float(margin), out_same))
    def classify(b, t, g):
        scale = 0.0
        fs = getattr(want[b], "frame_scale", None)
        if fs:
            scale = fs[min(t, len(fs) - 1)]
        if g < near_tie:
            rep.near_tie_frames.append((b, t, g))
        elif scale > 0 and g <= RESOLUTION_ULPS * float(np.spacing(np.float32(scale))) * 1.0001:
            rep.resolution_ties.append((b, t, g))
        else:
            rep.unexplained.append((b, t, g))

    if coupled and first:
        b0, t0_, g0, _ = min(first, key=lambda x: x[1])
        for b, t, g, _ in first:
            if (b, t) == (b0, t0_) or t < t0_ or g0 >= near_tie:
                classify(b, t, g)
            else:
                rep.cascade_streams.append(b)
    else:
        for b, t, g, _ in first:
            classify(b, t, g)
    return rep
