"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  ***parity unpinned***

A numpy (fp32) restatement of the transducer-search hot path of manyeyes/K2TransducerAsr. Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this module; the product (`k2transducerasr_b200/`, `libk2b200.so`) never does and has no
CPU fallback.

Why "parity unpinned": the reference ships no tests and no golden vectors for this path
(SURVEY.md section 4), it is 100 % C# over the un-vendored NuGet package
Microsoft.ML.OnnxRuntime 1.22.1 (ref K2TransducerAsr.csproj:14) running user-downloaded
decoder/joiner .onnx files, and neither a .NET toolchain nor onnxruntime exists in this image, so
the reference cannot be executed here. What IS in the reference tree — the control flow of the
search loops and every quirk of it — is restated below line by line with citations; the arithmetic
inside the ONNX graphs follows the k2-fsa/icefall export the reference README points at
(ref README.EN.md:285) and is marked [EXT]. `modified_beam_search` does not exist in the
reference at all (ref OnlineRecognizer.cs:19 has only the dead `maxActivePaths` argument); its
semantics are icefall's, also [EXT].

All arithmetic is float32. GEMMs go through numpy/OpenBLAS sgemm, which is what ONNX Runtime's
CPU provider (MLAS sgemm) amounts to; accumulation order therefore differs from any other
implementation by a few ulp, which is why token parity is judged with near-tie frames
(top-2 gap < 1e-4) listed separately (BASELINE.json north_star).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# operand rounding used to restate the tensor-core modes of the CUDA library on the CPU
# --------------------------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what cvt.rn.bf16.f32 does)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out.reshape(x.shape)


def _gemm_nt(a: np.ndarray, w: np.ndarray, prec: str) -> np.ndarray:
    """a [M,K] @ w[N,K].T under one of the library's arithmetic modes.
    fp32  : plain sgemm.
    bf16  : both operands rounded to bf16, products exact in fp32, fp32 accumulation.
    bf16x3: a = a_hi + a_lo, w = w_hi + w_lo (each bf16); a_hi*w_hi + a_hi*w_lo + a_lo*w_hi."""
    a = np.ascontiguousarray(a, dtype=F32)
    w = np.ascontiguousarray(w, dtype=F32)
    if prec == "fp32":
        return a @ w.T
    if prec == "bf16":
        return round_bf16(a) @ round_bf16(w).T
    if prec == "bf16x3":
        ah = round_bf16(a)
        al = round_bf16(a - ah)
        wh = round_bf16(w)
        wl = round_bf16(w - wh)
        return (ah @ wh.T) + ((ah @ wl.T) + (al @ wh.T))
    raise ValueError(prec)


# --------------------------------------------------------------------------------------------
# model math [EXT]: what decoder.onnx / joiner.onnx / the encoder_proj tail of encoder.onnx hold
# --------------------------------------------------------------------------------------------
@dataclass
class Model:
    """Weights in the layouts of include/k2b200.h::k2b_load_weights."""

    emb: np.ndarray          # [V,D]
    conv_w: np.ndarray       # [D,4,ctx]  Conv1d(D, D, kernel=ctx, groups=D/4, bias=False)
    dec_proj_w: np.ndarray   # [J,D]
    dec_proj_b: np.ndarray   # [J]
    out_w: np.ndarray        # [V,J]
    out_b: np.ndarray        # [V]
    enc_proj_w: Optional[np.ndarray] = None  # [J,E]
    enc_proj_b: Optional[np.ndarray] = None  # [J]
    blank_id: int = 0        # ref OfflineModel.cs:18
    sos_eos_id: int = 1      # ref OfflineModel.cs:19
    unk_id: int = 2          # ref OfflineModel.cs:20
    context_size: int = 2    # ref Model/OfflineCustomMetadata.cs:21
    neg_id_wrap: bool = False
    prec: str = "fp32"
    prec_enc: Optional[str] = None      # arithmetic of the encoder_proj GEMM only (None = prec)
    prec_joiner: Optional[str] = None   # arithmetic of the joiner output GEMM only (None = prec): restates the
                                        # library's tensor-core modes, whose decoder / encoder_proj stay fp32

    @property
    def V(self) -> int:
        return self.out_w.shape[0]

    @property
    def J(self) -> int:
        return self.out_w.shape[1]

    @classmethod
    def from_dict(cls, w: dict, **kw) -> "Model":
        return cls(emb=w["emb"], conv_w=w["conv_w"], dec_proj_w=w["dec_proj_w"], dec_proj_b=w["dec_proj_b"],
                   out_w=w["out_w"], out_b=w["out_b"], enc_proj_w=w.get("enc_proj_w"),
                   enc_proj_b=w.get("enc_proj_b"), **kw)


def decoder_conv(m: Model, y: np.ndarray) -> np.ndarray:
    """Embedding gather + grouped conv1d + ReLU -> [N,D].  [EXT] icefall Decoder.forward:
    emb = Emb[clamp(y,0)] * (y >= 0); Conv1d(kernel=ctx, groups=D/4) over the ctx positions."""
    y = np.asarray(y, dtype=np.int64).reshape(-1, m.context_size)
    n = y.shape[0]
    V, D = m.emb.shape
    if m.neg_id_wrap:
        e = m.emb[np.where(y < 0, y + V, y)]
    else:
        e = m.emb[np.clip(y, 0, None)] * (y >= 0)[..., None].astype(F32)
    e = e.astype(F32)                                   # [N,ctx,D]
    g = e.reshape(n, m.context_size, D // 4, 4)          # input channels of group o//4
    wc = m.conv_w.reshape(D // 4, 4, 4, m.context_size)  # [group, out-in-group, in-in-group, k]
    c = np.einsum("nkgi,goik->ngo", g, wc, dtype=F32, optimize=False).reshape(n, D)
    return np.maximum(c, F32(0)).astype(F32)


def decoder(m: Model, y: Optional[np.ndarray], n: Optional[int] = None) -> np.ndarray:
    """DecoderProj (ref OfflineProjOfTransducer.cs:93-123): y == None -> n x {-1, blank}
    (ref :97-110). Returns decoder_proj(relu(conv(emb(y)))) [N,J]. [EXT] for the math."""
    if y is None:
        y = np.tile(np.array([-1, m.blank_id], dtype=np.int64), (n or 1, 1))
    r = decoder_conv(m, y)
    return (_gemm_nt(r, m.dec_proj_w, m.prec) + m.dec_proj_b).astype(F32)


def joiner(m: Model, enc: np.ndarray, dec: np.ndarray) -> np.ndarray:
    """JoinerProj (ref OfflineProjOfTransducer.cs:125-152): logits = out_linear(tanh(enc + dec)),
    no softmax. [EXT] for the math."""
    x = np.tanh((np.asarray(enc, F32).reshape(-1, m.J) + np.asarray(dec, F32).reshape(-1, m.J)).astype(F32)).astype(F32)
    return (_gemm_nt(x, m.out_w, m.prec_joiner or m.prec) + m.out_b).astype(F32)


def encoder_proj(m: Model, raw: np.ndarray) -> np.ndarray:
    """The Linear E->J that upstream folds into encoder.onnx [EXT]; the reference only ever sees its
    output (ref OfflineProjOfTransducer.cs:83)."""
    E = m.enc_proj_w.shape[1]
    shp = raw.shape[:-1]
    out = _gemm_nt(np.asarray(raw, F32).reshape(-1, E), m.enc_proj_w, m.prec_enc or m.prec) + m.enc_proj_b
    return out.astype(F32).reshape(*shp, m.J)


# --------------------------------------------------------------------------------------------
# the two argmax rules of the reference
# --------------------------------------------------------------------------------------------
def argmax_hi(logits: np.ndarray) -> np.ndarray:
    """Transducer argmax, ref OfflineRecognizer.cs:150-154 / :236-240, OnlineRecognizer.cs:159-163:
        token = 0; for k in 1..V-1: token = logits[token] > logits[k] ? token : k
    => ties AND NaN comparisons go to the larger index (Q1). Returns int64 [N]."""
    lg = np.asarray(logits, F32)
    if lg.ndim == 1:
        lg = lg[None, :]
    n, V = lg.shape
    out = (V - 1 - np.argmax(lg[:, ::-1], axis=1)).astype(np.int64)
    bad = np.isnan(lg).any(axis=1)
    for r in np.nonzero(bad)[0]:          # literal fold for rows holding NaN
        tok = 0
        row = lg[r]
        for k in range(1, V):
            tok = tok if row[tok] > row[k] else k
        out[r] = tok
    return out


def argmax_lo(row: np.ndarray) -> int:
    """CTC argmax, ref OfflineRecognizer.cs:335-336: IndexOf(slice.Max()) - first index of the
    maximum (Q2). .NET's Max<float> orders NaN below every number and IndexOf uses Equals
    (NaN equals NaN, -0 equals +0): an all-NaN frame gives index 0."""
    row = np.asarray(row, F32)
    nan = np.isnan(row)
    if not nan.any():
        return int(np.argmax(row))
    if nan.all():
        return 0
    v = np.where(nan, -np.inf, row)
    mx = v.max()
    return int(np.nonzero((row == mx) & ~nan)[0][0])


def top2_gap(logits: np.ndarray) -> np.ndarray:
    """best - second best per row (the near-tie diagnostic of the north star)."""
    lg = np.asarray(logits, F32)
    if lg.shape[1] < 2:
        return np.full(lg.shape[0], np.inf, F32)
    part = np.partition(lg, lg.shape[1] - 2, axis=1)
    return (part[:, -1] - part[:, -2]).astype(F32)


# --------------------------------------------------------------------------------------------
# results
# --------------------------------------------------------------------------------------------
@dataclass
class StreamResult:
    tokens: List[int]              # exactly what the reference leaves in stream.Tokens
    timestamps: List[int]          # what this call appends to stream.Timestamps
    appended: List[int] = field(default_factory=list)   # symbols emitted by this call only
    num_trailing_blank: int = 0
    score: float = 0.0             # mbs: un-normalised log-prob of the chosen hypothesis
    min_gap: float = float("inf")  # smallest decision margin met while decoding this stream
    hyp: Optional[List[int]] = None  # online: stream.Hyp after the call
    frame_gap: List[float] = field(default_factory=list)   # decision margin of every decoded frame (greedy: top-2 logit gap;
                                                           # mbs: smallest adjacent gap among the stream's top-(beam+1) scores)
    history: Optional[List[List[tuple]]] = None   # mbs: per frame, the surviving hypotheses in insertion order as
                                                  # (parent slot in the previous frame's list, appended token or -1)
    final_gap: float = float("inf")               # mbs: margin of the final length-normalised pick
    frame_scale: List[float] = field(default_factory=list)   # mbs: |best hypothesis score| of every frame - the float32 spacing
                                                             # at that magnitude bounds what ANY fp32 implementation can resolve


# --------------------------------------------------------------------------------------------
# A.1 offline single-stream greedy  (ref OfflineRecognizer.cs:93-187)
# --------------------------------------------------------------------------------------------
def greedy_search_single(m: Model, enc: np.ndarray, max_sym_per_utt: int = 1000,
                         max_sym_per_frame: int = 1) -> StreamResult:
    """enc [T,J] projected frames. The loop of ref :127-179 verbatim: a frame is re-evaluated with the
    refreshed decoder output until it yields blank / unk or `max_sym_per_frame` symbols were emitted on it
    (the reference fixes max_sym_per_frame = 1, ref :19 => one joiner evaluation per frame); emit unless
    y in {blank, unk} (ref :161); decoder refreshed on emission (ref :165-170); Tokens start {-1, blank}
    (ref :115-117); timestamps = frame index (ref :164). frame_gap[t] = the smallest top-2 margin of the
    evaluations made on frame t."""
    enc = np.asarray(enc, F32).reshape(-1, m.J)
    T = enc.shape[0]
    toks = [-1, m.blank_id]
    ts: List[int] = []
    d = decoder(m, np.array([[-1, m.blank_id]], np.int64))
    t, n_sym, sym_per_frame = 0, 0, 0
    fgap = [float("inf")] * T
    while t < T and n_sym < max_sym_per_utt:
        if sym_per_frame >= max_sym_per_frame:      # ref :129-134
            sym_per_frame = 0
            t += 1
            continue
        lg = joiner(m, enc[t:t + 1], d)
        y = int(argmax_hi(lg)[0])
        fgap[t] = min(fgap[t], float(top2_gap(lg)[0]))
        if y != m.blank_id and y != m.unk_id:
            toks.append(y)
            ts.append(t)
            d = decoder(m, np.array([toks[-m.context_size:]], np.int64))
            n_sym += 1
            sym_per_frame += 1
        else:                                       # ref :174-178
            sym_per_frame = 0
            t += 1
    return StreamResult(tokens=toks, timestamps=ts, appended=toks[2:], min_gap=min(fgap, default=float("inf")),
                        frame_gap=fgap)


# --------------------------------------------------------------------------------------------
# A.2 offline batch greedy (ref OfflineRecognizer.cs:189-303) + the per-stream-consistent mode
# --------------------------------------------------------------------------------------------
def greedy_search_batch(m: Model, enc: np.ndarray, compat: bool = True) -> List[StreamResult]:
    """enc [B,T,J]. compat=True restates the reference including Q5 (2B blank seed, ref :250-267),
    Q6 (decoder re-run for ALL streams when ANY emits, from the token-list tails, ref :278-286) and
    Q7 (no length masking). compat=False: every stream behaves as greedy_search_single."""
    enc = np.asarray(enc, F32)
    B, T, _ = enc.shape
    blank = m.blank_id
    d = decoder(m, None, B)
    if compat:
        toks = [[blank] * (2 * B) for _ in range(B)]
        ts = [[0] * (2 * B) for _ in range(B)]
    else:
        toks = [[-1, blank] for _ in range(B)]
        ts = [[] for _ in range(B)]
    n0 = len(toks[0]) if B else 0
    gap = np.full(B, np.inf, F32)
    fgap = np.full((B, T), np.inf, F32)
    for t in range(T):
        lg = joiner(m, enc[:, t, :], d)
        y = argmax_hi(lg)
        fgap[:, t] = top2_gap(lg)
        gap = np.minimum(gap, fgap[:, t])
        emitted = False
        for b in range(B):
            if y[b] != blank and y[b] != m.unk_id:
                toks[b].append(int(y[b]))
                ts[b].append(t)
                emitted = True
        if emitted:
            d = decoder(m, np.array([tk[-m.context_size:] for tk in toks], np.int64))
    if T == 0 and compat:
        # Q13: the reference leaves tokens[m] null here; we define T == 0 as "nothing appended".
        toks = [[] for _ in range(B)]
        ts = [[] for _ in range(B)]
        n0 = 0
    return [StreamResult(tokens=toks[b], timestamps=ts[b], appended=toks[b][n0:], min_gap=float(gap[b]),
                         frame_gap=[float(g) for g in fgap[b]]) for b in range(B)]


# --------------------------------------------------------------------------------------------
# A.3 online batch greedy, one chunk (ref OnlineRecognizer.cs:85-219)
# --------------------------------------------------------------------------------------------
def greedy_search_online_chunk(m: Model, enc: np.ndarray, hyps: Sequence[Sequence[int]],
                               tokens: Sequence[List[int]]) -> List[StreamResult]:
    """enc [B,T',J]; hyps[b] = stream.Hyp (ref :109), tokens[b] = stream.Tokens (ref :111), both
    initially {blank, blank} (ref OnlineStream.cs:44-45). Emission mask is {blank, unk, 1}
    (ref :181, the literal 1); timestamps are chunk-local (ref :184); Hyp out = last ctx tokens
    (ref :208)."""
    enc = np.asarray(enc, F32)
    B, T, _ = enc.shape
    toks = [list(tk) for tk in tokens]
    n0 = [len(tk) for tk in toks]
    ts: List[List[int]] = [[] for _ in range(B)]
    d = decoder(m, np.array([list(h) for h in hyps], np.int64).reshape(B, m.context_size))
    gap = np.full(B, np.inf, F32)
    fgap = np.full((B, T), np.inf, F32)
    for t in range(T):
        lg = joiner(m, enc[:, t, :], d)
        y = argmax_hi(lg)
        fgap[:, t] = top2_gap(lg)
        gap = np.minimum(gap, fgap[:, t])
        emitted = False
        for b in range(B):
            if y[b] != m.blank_id and y[b] != m.unk_id and y[b] != 1:
                toks[b].append(int(y[b]))
                ts[b].append(t)
                emitted = True
        if emitted:
            d = decoder(m, np.array([tk[-m.context_size:] for tk in toks], np.int64))
    return [StreamResult(tokens=toks[b], timestamps=ts[b], appended=toks[b][n0[b]:], min_gap=float(gap[b]),
                         hyp=toks[b][-m.context_size:], frame_gap=[float(g) for g in fgap[b]]) for b in range(B)]


# --------------------------------------------------------------------------------------------
# A.4 CTC greedy (ref OfflineRecognizer.cs:305-430, OnlineRecognizer.cs:220-319)
# --------------------------------------------------------------------------------------------
def ctc_greedy_search(logp: np.ndarray, blank: int = 0, frame_offset: Optional[Sequence[int]] = None,
                      trailing_blank: Optional[Sequence[int]] = None,
                      prev: Optional[Sequence[int]] = None) -> List[StreamResult]:
    """logp [B,T,V]. Per frame y = first index of the max (ref :335); trailing-blank counter
    (ref :337-344); emit when y != blank and y != prev_id (ref :346-350); prev_id starts at -1 on
    every call (ref :332) unless `prev` carries it across calls (our option, Q10). `appended` holds
    this call's symbols; the caller prepends the stream's own Tokens seed."""
    logp = np.asarray(logp, F32)
    B, T, V = logp.shape
    out = []
    for b in range(B):
        off = int(frame_offset[b]) if frame_offset is not None else 0
        ntb = int(trailing_blank[b]) if trailing_blank is not None else 0
        prev_id = int(prev[b]) if prev is not None else -1
        toks: List[int] = []
        ts: List[int] = []
        gap = float("inf")
        rows = logp[b]
        nanrow = np.isnan(rows).any(axis=1) if T else np.zeros(0, bool)
        ys = np.argmax(rows, axis=1) if T else np.zeros(0, np.int64)
        for t in range(T):
            y = argmax_lo(rows[t]) if nanrow[t] else int(ys[t])
            ntb = ntb + 1 if y == blank else 0
            if y != blank and y != prev_id:
                toks.append(y)
                ts.append(t + off)
            prev_id = y
        fg: List[float] = []
        if T and V > 1 and not nanrow.any():
            g2 = top2_gap(rows)
            gap = float(g2.min())
            fg = [float(x) for x in g2]
        r = StreamResult(tokens=toks, timestamps=ts, appended=toks, num_trailing_blank=ntb, min_gap=gap, frame_gap=fg)
        r.hyp = [prev_id]
        out.append(r)
    return out


# --------------------------------------------------------------------------------------------
# A.5 modified_beam_search  [EXT - icefall semantics; absent from the reference]
# --------------------------------------------------------------------------------------------
def log_softmax(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, F32)
    mx = x.max(axis=-1, keepdims=True)
    z = (x - mx).astype(F32)
    lse = np.log(np.exp(z).sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
    return (z - lse).astype(F32)


def logaddexp32(a: float, b: float) -> np.float32:
    a, b = F32(a), F32(b)
    mx, mn = (a, b) if a >= b else (b, a)
    return F32(mx + np.log1p(np.exp(F32(mn - mx)), dtype=F32))


@dataclass
class _Hyp:
    ys: List[int]
    lp: np.float32
    ts: List[int]
    cs: int = 0          # context-graph (hot word) state


def mbs_seed(m: Model, B: int, hyp: Optional[Sequence[Sequence[int]]] = None) -> List[List[_Hyp]]:
    """Initial beam of B streams: one hypothesis {ys = [-1]*(ctx-1) + [blank], lp = 0} (offline, the seed of
    ref OfflineRecognizer.cs:105), or ys = stream.Hyp (online: {blank, blank}, ref OnlineStream.cs:44-45)."""
    if hyp is None:
        return [[_Hyp([-1] * (m.context_size - 1) + [m.blank_id], F32(0), [])] for _ in range(B)]
    return [[_Hyp([int(x) for x in hyp[b]], F32(0), [])] for b in range(B)]


def modified_beam_search(m: Model, enc: np.ndarray, beam: int = 4, init: Optional[List[List[_Hyp]]] = None,
                         frame_offset: Optional[Sequence[int]] = None, extra_mask: Optional[int] = None,
                         return_state: bool = False, context_graph=None, finalize: bool = True):
    """enc [B,T,J]. Per stream keep <= beam hypotheses, seeded {ys=[-1]*(ctx-1)+[blank], lp=0}.
    Per frame: decoder on every live hypothesis' last ctx tokens -> joiner with the stream's frame
    -> log_softmax -> + hyp.lp -> top-`beam` over the stream's flattened [n_hyps*V] scores ->
    extend (ys unchanged for blank/unk) -> insert keyed by ys, log-add on collision, first-inserted
    hypothesis keeps its timestamps. Result: argmax lp/len(ys) (len counts the seeds), first
    maximum in insertion order.
    Stated choices nothing pins: top-k order is value-descending then flat-index-DEscending
    (consistent with Q1: beam=1 equals greedy_search_single), temperature 1, no blank penalty.
    Streaming (the dead maxActivePaths of ref OnlineRecognizer.cs:19 brought to life): `init` = the beams a
    previous chunk returned (return_state=True), `frame_offset[b]` = frames of stream b decoded before this
    chunk (timestamps are utterance-absolute), `extra_mask` = a third non-emitting id (the literal 1 of
    ref OnlineRecognizer.cs:181). Decoding an utterance chunk by chunk equals decoding it whole.
    Hot words (`context_graph` = k2transducerasr_b200.hotwords.ContextGraph: dense next / delta / residual tables, semantics in
    that module): the boost delta[state, token] is added to an EXTENDED hypothesis after the top-`beam` selection (icefall's order
    [EXT]) and the hypothesis moves to next[state, token]; with `finalize` the residual of the final state is subtracted before the
    hypotheses are compared (offline); streaming results "so far" keep it (finalize=False)."""
    enc = np.asarray(enc, F32)
    B, T, _ = enc.shape
    V = m.V
    hyps: List[List[_Hyp]] = mbs_seed(m, B) if init is None else [[_Hyp(list(h.ys), F32(h.lp), list(h.ts), h.cs) for h in hs] for hs in init]
    foff = [0] * B if frame_offset is None else [int(x) for x in frame_offset]
    gap = np.full(B, np.inf, np.float64)
    fgap = np.full((B, T), np.inf, np.float64)
    fscale = np.zeros((B, T), np.float64)
    history: List[List[List[tuple]]] = [[] for _ in range(B)]
    for t in range(T):
        counts = [len(h) for h in hyps]
        ctx = np.array([h.ys[-m.context_size:] for hs in hyps for h in hs], np.int64)
        d = decoder(m, ctx)
        rows = np.repeat(np.arange(B), counts)
        lp = log_softmax(joiner(m, enc[rows, t, :], d))
        prev = np.array([h.lp for hs in hyps for h in hs], F32)
        lp = (lp + prev[:, None]).astype(F32)
        off = 0
        for b in range(B):
            n = counts[b]
            flat = lp[off:off + n].reshape(-1)
            off += n
            k = min(beam, flat.size)
            # value desc, flat index desc; exact top-(k+1) without sorting all n*V scores
            kk = min(k + 1, flat.size)
            thr = np.partition(flat, flat.size - kk)[flat.size - kk]
            cand = np.nonzero(flat >= thr)[0]
            order = cand[np.lexsort((-cand, -flat[cand].astype(np.float64)))]
            top = order[:k]
            # decision margins: the beam boundary AND the order inside the beam (it decides which
            # of two colliding hypotheses is inserted first and so keeps its timestamps)
            head = flat[order[:min(k + 1, flat.size)]].astype(np.float64)
            fscale[b, t] = abs(float(head[0])) if head.size else 0.0
            if head.size > 1:
                fgap[b, t] = float(np.min(head[:-1] - head[1:]))
                gap[b] = min(gap[b], fgap[b, t])
            new: List[_Hyp] = []
            index = {}
            rec: List[tuple] = []
            for fi in top:
                par = int(fi) // V
                h = hyps[b][par]
                tok = int(fi) % V
                ys, tss, cs = h.ys, h.ts, h.cs
                emit = tok != m.blank_id and tok != m.unk_id and tok != extra_mask
                new_lp = F32(flat[fi])
                if emit:
                    ys = ys + [tok]
                    tss = tss + [t + foff[b]]
                    if context_graph is not None:
                        new_lp = F32(new_lp + F32(context_graph.delta[cs, tok]))
                        cs = int(context_graph.next[cs, tok])
                key = tuple(ys)
                if key in index:
                    o = new[index[key]]
                    o.lp = logaddexp32(o.lp, new_lp)
                else:
                    index[key] = len(new)
                    new.append(_Hyp(list(ys), new_lp, list(tss), cs))
                    rec.append((par, tok if emit else -1))
            hyps[b] = new
            history[b].append(rec)
    out = []
    for b in range(B):
        fin_lp = [F32(h.lp) - (F32(context_graph.residual[h.cs]) if (context_graph is not None and finalize) else F32(0)) for h in hyps[b]]
        norm = [F32(v) / F32(len(h.ys)) for v, h in zip(fin_lp, hyps[b])]
        bi = 0
        for i in range(1, len(norm)):
            if norm[i] > norm[bi]:
                bi = i
        fin = float("inf")
        if len(norm) > 1:
            srt = sorted(float(x) for x in norm)
            fin = srt[-1] - srt[-2]
            gap[b] = min(gap[b], fin)
        h = hyps[b][bi]
        out.append(StreamResult(tokens=h.ys, timestamps=h.ts, appended=h.ys[m.context_size:], score=float(fin_lp[bi]),
                                min_gap=float(gap[b]), hyp=h.ys[-m.context_size:], frame_gap=[float(g) for g in fgap[b]],
                                history=history[b], final_gap=fin, frame_scale=[float(x) for x in fscale[b]]))
    return (out, hyps) if return_state else out


def ragged(search, m: Model, enc: np.ndarray, lens: Sequence[int], *args, **kw) -> List[StreamResult]:
    """Ragged batch = every stream decoded alone over its own frames [0, lens[b]) (streams are independent in
    modified_beam_search and in the per-stream greedy loop, ref OfflineRecognizer.cs:93-187). The reference itself never
    consumes encoder_out_lens (ref OfflineProjOfTransducer.cs:84, Q7: padding is decoded like speech); this is the
    semantics of k2b_set_encoder_out_lens. `search` is modified_beam_search or greedy_search_batch (compat=False)."""
    enc = np.asarray(enc, F32)
    out: List[StreamResult] = []
    for b, n in enumerate(lens):
        out.extend(search(m, enc[b:b + 1, :int(n)], *args, **kw))
    return out


def stack_states(items: Sequence[Sequence[np.ndarray]], axis_len: Sequence[int]) -> List[np.ndarray]:
    """ref OnlineProjOfZipformer2.cs:144-362 (and the Zipformer / Lstm / Conformer siblings): per cache tensor i, with A = its
    "axisnum", stacked_i[(x*B + n)*A + a] = items[n][i][x*A + a] - the Array.Copy loops, e.g. :237-246 for cached_key."""
    B = len(items)
    out = []
    for i, A in enumerate(axis_len):
        L = items[0][i].size
        st = np.zeros(L * B, F32)
        for x in range(L // A):
            for n in range(B):
                st[(x * B + n) * A:(x * B + n + 1) * A] = items[n][i][x * A:(x + 1) * A]
        out.append(st)
    return out


def unstack_states(stacked: Sequence[np.ndarray], B: int, axis_len: Sequence[int]) -> List[List[np.ndarray]]:
    """ref OnlineProjOfZipformer2.cs:363-489: item_{n,i}[k*A + a] = stacked_i[(B*k + n)*A + a] (e.g. :399-404)."""
    items = [[None] * len(axis_len) for _ in range(B)]
    for i, A in enumerate(axis_len):
        L = stacked[i].size // B
        for n in range(B):
            it = np.zeros(L, F32)
            for k in range(L // A):
                it[k * A:(k + 1) * A] = stacked[i][(B * k + n) * A:(B * k + n + 1) * A]
            items[n][i] = it
    return items
