"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs,
against the committed golden fixtures, and - at BASELINE.json's full sizes - through size-independent
properties. Bars: bit-exact for the CTC path (integer / compare work); for the transducer paths token
sequences identical except streams where the oracle itself met a near tie (gap < 1e-4), hypothesis
log-probs within 1e-3 absolute (fp32)."""
from pathlib import Path

import numpy as np
import pytest

from k2transducerasr_b200 import _native, synth
from k2transducerasr_b200 import proj as P
from k2transducerasr_b200 import recognizer as R
from oracle import k2_oracle as O
from tests.helpers import MID, NEAR_TIE, SCORE_TOL, SMALL, compare_streams, model_and_weights

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "small_v53.npz")


def _unragged(a, n):
    return [a[i, :n[i]].tolist() for i in range(len(n))]


def make_handle(dims, w, **kw):
    h = _native.Handle(vocab_size=dims.vocab_size, joiner_dim=dims.joiner_dim, decoder_dim=dims.decoder_dim,
                       encoder_dim=dims.encoder_dim, **kw)
    if w is not None:
        h.load_weights(w)
    return h


@pytest.fixture(scope="module")
def small(built_lib):
    m, w = model_and_weights(SMALL, blank_bias=0.6)
    h = make_handle(SMALL, w)
    yield m, w, h
    h.close()


@pytest.fixture(scope="module")
def mid(built_lib):
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make_handle(MID, w)
    yield m, w, h
    h.close()


# ---- fine-grained seam -------------------------------------------------------------------------------------
def test_native_library_is_loaded(built_lib):
    import ctypes
    maps = open("/proc/self/maps").read()
    _native.lib()
    maps = open("/proc/self/maps").read()
    assert "libk2b200.so" in maps


def test_golden_model_math(small):
    m, w, h = small
    np.testing.assert_allclose(h.encoder_proj(GOLD["raw"]), GOLD["enc"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(h.decoder_proj(GOLD["y"]), GOLD["dec"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(h.joiner_proj(GOLD["enc"][:, 0, :], GOLD["dec"]), GOLD["logits"], atol=1e-5, rtol=0)


def test_decoder_null_input_and_negative_ids(small):
    m, w, h = small
    np.testing.assert_allclose(h.decoder_proj(None, 3), O.decoder(m, None, 3), atol=1e-5)     # ref :97-110
    hw = make_handle(SMALL, w, neg_id_mode=_native.NEGID_WRAP)
    mw = O.Model.from_dict(w, neg_id_wrap=True)
    y = np.array([[-1, 0], [-1, 5]], np.int64)
    np.testing.assert_allclose(hw.decoder_proj(y), O.decoder(mw, y), atol=1e-5)
    assert np.abs(hw.decoder_proj(y) - h.decoder_proj(y)).max() > 1e-3
    hw.close()


@pytest.mark.parametrize("n", [1, 3, 64, 65, 300])
def test_seam_shapes_mid(mid, n):
    m, w, h = mid
    rng = np.random.default_rng(n)
    y = rng.integers(-1, MID.vocab_size, size=(n, 2)).astype(np.int64)
    y[:, 1] = np.abs(y[:, 1])
    dec = h.decoder_proj(y)
    np.testing.assert_allclose(dec, O.decoder(m, y), atol=2e-5, rtol=0)
    raw = rng.standard_normal((n, MID.encoder_dim), dtype=np.float32)
    enc = h.encoder_proj(raw)
    np.testing.assert_allclose(enc, O.encoder_proj(m, raw), atol=2e-5, rtol=0)
    np.testing.assert_allclose(h.joiner_proj(enc, dec), O.joiner(m, enc, dec), atol=2e-5, rtol=0)


# ---- greedy ----------------------------------------------------------------------------------------------------
def test_golden_greedy_and_beam(small):
    m, w, h = small
    enc = GOLD["enc"]
    toks, tss = h.greedy_offline(enc[:1], _native.GREEDY_SINGLE)
    assert [-1, 0] + toks[0] == GOLD["single_tokens"].tolist() and tss[0] == GOLD["single_ts"].tolist()
    B = enc.shape[0]
    toks, tss = h.greedy_offline(enc, _native.GREEDY_BATCH_COMPAT)
    assert [[0] * (2 * B) + t for t in toks] == _unragged(GOLD["compat_tokens"], GOLD["compat_ntok"])
    assert [[0] * (2 * B) + t for t in tss] == _unragged(GOLD["compat_ts"], GOLD["compat_nts"])
    toks, tss = h.greedy_offline(enc, _native.GREEDY_PER_STREAM)
    assert [[-1, 0] + t for t in toks] == _unragged(GOLD["perstream_tokens"], GOLD["perstream_ntok"])
    for K in (4, 2):
        toks, tss, sc = h.modified_beam_search(enc, K)
        assert [[-1, 0] + t for t in toks] == _unragged(GOLD[f"mbs{K}_tokens"], GOLD[f"mbs{K}_ntok"])
        assert tss == _unragged(GOLD[f"mbs{K}_ts"], GOLD[f"mbs{K}_nts"])
        np.testing.assert_allclose(sc, GOLD[f"mbs{K}_score"], atol=SCORE_TOL)


def test_golden_online_chunks(small):
    m, w, h = small
    enc = GOLD["enc"]
    B = enc.shape[0]
    hyp = np.zeros((B, 2), np.int64)
    toks = [[0, 0] for _ in range(B)]
    tss = [[] for _ in range(B)]
    for c in range(3):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(enc[:, 8 * c:8 * c + 8]), hyp)
        for b in range(B):
            toks[b] += t[b]
            tss[b] += s[b]
    assert toks == _unragged(GOLD["online_tokens"], GOLD["online_ntok"])
    assert tss == _unragged(GOLD["online_ts"], GOLD["online_nts"])
    assert hyp.tolist() == GOLD["online_hyp"].tolist()


@pytest.mark.parametrize("raw_regime", [False, True])
def test_greedy_modes_vs_oracle_mid(mid, raw_regime):
    m, w, h = mid
    if raw_regime:                       # no blank bias: nearly every frame emits (worst case for the decoder)
        m, w = model_and_weights(MID, blank_bias=0.0)
        h = make_handle(MID, w)
    raw = synth.make_frames(9, 40, MID.encoder_dim, 31)
    enc = O.encoder_proj(m, raw)
    t, s = h.greedy_offline(enc[:1], _native.GREEDY_SINGLE)
    compare_streams(t, s, [O.greedy_search_single(m, enc[0])], "single")
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT)            # raw frames: encoder_proj on device
    compare_streams(t, s, O.greedy_search_batch(m, enc, True), "batch_compat")
    t, s = h.greedy_offline(enc, _native.GREEDY_PER_STREAM)
    compare_streams(t, s, O.greedy_search_batch(m, enc, False), "per_stream")
    if raw_regime:
        h.close()


def test_q6_batch_composition_changes_compat_but_not_per_stream(mid):
    m, w, h = mid
    enc = O.encoder_proj(m, synth.make_frames(6, 30, MID.encoder_dim, 32))
    alone, _ = h.greedy_offline(enc[2:3], _native.GREEDY_PER_STREAM)
    inside, _ = h.greedy_offline(enc, _native.GREEDY_PER_STREAM)
    assert alone[0] == inside[2]


def test_q1_q3_crafted_rows(small):
    """Ties go to the larger index; unk is never emitted offline; id 1 is masked online only."""
    m, w, h = small
    w2 = {k: (None if v is None else v.copy()) for k, v in w.items()}
    w2["out_w"][:] = 0.0
    w2["out_b"][:] = -1.0
    w2["out_b"][[7, 19, 33]] = 2.0                         # exact 3-way tie on every frame -> 33
    h2 = make_handle(SMALL, w2)
    enc = synth.make_frames(3, 5, SMALL.joiner_dim, 1)
    toks, _ = h2.greedy_offline(enc, _native.GREEDY_PER_STREAM)
    assert toks == [[33] * 5] * 3
    w2["out_b"][[7, 19, 33]] = -1.0
    w2["out_b"][2] = 5.0                                   # unk wins
    h2.load_weights(w2)
    assert h2.greedy_offline(enc, _native.GREEDY_BATCH_COMPAT)[0] == [[]] * 3
    w2["out_b"][2] = -1.0
    w2["out_b"][1] = 5.0                                   # id 1 wins
    h2.load_weights(w2)
    assert h2.greedy_offline(enc, _native.GREEDY_PER_STREAM)[0] == [[1] * 5] * 3
    t, s, hyp = h2.greedy_online_chunk(enc, np.zeros((3, 2), np.int64))
    assert t == [[]] * 3 and hyp.tolist() == [[0, 0]] * 3
    w2["out_b"][1] = np.nan                                # NaN logit: the fold lands on the next index (Q1)
    h2.load_weights(w2)
    toks, _ = h2.greedy_offline(enc[:1, :1], _native.GREEDY_PER_STREAM)
    mm = O.Model.from_dict(w2)
    want = O.greedy_search_batch(mm, enc[:1, :1], False)
    assert toks[0] == want[0].appended
    h2.close()


def test_online_chunks_vs_oracle_mid(mid):
    m, w, h = mid
    B, Tc, C = 12, 8, 4
    raw = synth.make_frames(B, Tc * C, MID.encoder_dim, 33)
    enc = O.encoder_proj(m, raw)
    hyp = np.zeros((B, 2), np.int64)
    ohyp, otoks = [[0, 0]] * B, [[0, 0]] * B
    for c in range(C):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp)
        res = O.greedy_search_online_chunk(m, enc[:, Tc * c:Tc * c + Tc], ohyp, otoks)
        ex = compare_streams(t, s, res, f"online chunk {c}")
        if ex:
            pytest.skip("near-tie divergence: later chunks are not comparable")
        ohyp, otoks = [r.hyp for r in res], [r.tokens for r in res]
        assert hyp.tolist() == ohyp


def test_empty_inputs(mid):
    m, w, h = mid
    t, s = h.greedy_offline(np.zeros((3, 0, MID.joiner_dim), np.float32), _native.GREEDY_BATCH_COMPAT)
    assert t == [[], [], []]
    t, s, sc = h.modified_beam_search(np.zeros((2, 0, MID.joiner_dim), np.float32), 4)
    assert t == [[], []] and sc.tolist() == [0.0, 0.0]
    t, s, tb, _ = h.ctc_greedy(np.zeros((2, 0, 17), np.float32), trailing_blank=np.array([4, 0], np.int32))
    assert t == [[], []]
    with pytest.raises(_native.K2bError):
        h.greedy_offline(np.zeros((2, 4, MID.joiner_dim), np.float32), _native.GREEDY_SINGLE)   # B must be 1


# ---- modified_beam_search ------------------------------------------------------------------------------------
@pytest.mark.parametrize("beam", [1, 4, 8])
def test_beam_search_vs_oracle_mid(mid, beam):
    m, w, h = mid
    raw = synth.make_frames(10, 40, MID.encoder_dim, 34)
    enc = O.encoder_proj(m, raw)
    want = O.modified_beam_search(m, enc, beam)
    t, s, sc = h.modified_beam_search(raw, beam)
    bp = h.debug_backpointers(10, 40, beam)
    ex = compare_streams(t, s, want, f"mbs beam={beam}", bp=bp, scores=sc, allow_frac=0.1)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
    if beam == 1:
        g, gs = h.greedy_offline(enc, _native.GREEDY_PER_STREAM)
        assert [x for i, x in enumerate(g) if i not in ex] == [x for i, x in enumerate(t) if i not in ex]


def test_beam_merge_cases(small):
    m, w, h = small
    w2 = {k: (None if v is None else v.copy()) for k, v in w.items()}
    w2["out_b"][0] += 6.0
    w2["out_b"][2] += 6.0
    h2 = make_handle(SMALL, w2)
    m2 = O.Model.from_dict(w2)
    enc = synth.make_frames(3, 6, SMALL.joiner_dim, 8)
    want = O.modified_beam_search(m2, enc, 2)
    t, s, sc = h2.modified_beam_search(enc, 2)
    compare_streams(t, s, want, "merge blank/unk")
    np.testing.assert_allclose(sc, [r.score for r in want], atol=SCORE_TOL)
    h2.close()


# ---- CTC -------------------------------------------------------------------------------------------------------
def test_golden_ctc(small):
    m, w, h = small
    t, s, tb, _ = h.ctc_greedy(GOLD["ctc_logp"], 0, frame_offset=[0, 5, 0, 0], trailing_blank=np.array([0, 2, 0, 0]))
    assert t == _unragged(GOLD["ctc_tokens"], GOLD["ctc_ntok"])
    assert s == _unragged(GOLD["ctc_ts"], GOLD["ctc_nts"])
    assert tb.tolist() == GOLD["ctc_trailing"].tolist()


@pytest.mark.parametrize("form", [1, 0])      # k2b_set_option("ctc_one_kernel"): tickets in one kernel / frames + collapse kernels
@pytest.mark.parametrize("B,T,V", [(1, 1, 1), (3, 33, 2), (5, 64, 37), (4, 100, 2000), (3, 70, 5537), (2, 250, 501), (9, 40, 2051)])
def test_ctc_bit_exact_vs_oracle(small, B, T, V, form):
    m, w, h = small
    h.set_option("ctc_one_kernel", form)
    lp = synth.make_ctc_logp(B, T, V, 1000 + V, blank_bias=2.0 if V > 2 else 0.0)
    fo = np.arange(B, dtype=np.int32) * 3
    tb0 = np.arange(B, dtype=np.int32)
    want = O.ctc_greedy_search(lp, 0, frame_offset=fo, trailing_blank=tb0)
    t, s, tb, _ = h.ctc_greedy(lp, 0, frame_offset=fo, trailing_blank=tb0)
    assert t == [r.appended for r in want]
    assert s == [r.timestamps for r in want]
    assert tb.tolist() == [r.num_trailing_blank for r in want]
    h.set_option("ctc_one_kernel", -1)


@pytest.mark.parametrize("form", [1, 0])
def test_ctc_ties_nan_and_chunk_carry(small, form):
    m, w, h = small
    h.set_option("ctc_one_kernel", form)
    lp = synth.make_ctc_logp(2, 40, 101, 5, blank_bias=1.0)
    lp[0, 3, :] = -2.0                                    # whole-frame tie -> index 0 (blank)
    lp[0, 4, 10:] = -9.0; lp[0, 4, :10] = -8.0; lp[0, 4, 6] = lp[0, 4, 9] = -1.0    # tie -> 6
    lp[1, 5, :] = np.nan                                  # all NaN -> 0
    lp[1, 6, 0:50] = np.nan                               # NaN ordered below numbers
    want = O.ctc_greedy_search(lp)
    t, s, _, _ = h.ctc_greedy(lp)
    assert t == [r.appended for r in want] and s == [r.timestamps for r in want]
    # Q10: per-chunk reset vs carried prev
    a = h.ctc_greedy(np.ascontiguousarray(lp[:, :20]), prev=np.full(2, -1, np.int64), trailing_blank=np.zeros(2, np.int32))
    b = h.ctc_greedy(np.ascontiguousarray(lp[:, 20:]), prev=a[3], trailing_blank=a[2], frame_offset=[20, 20])
    assert [x + y for x, y in zip(a[0], b[0])] == [r.appended for r in want]
    assert [x + y for x, y in zip(a[1], b[1])] == [r.timestamps for r in want]
    assert b[2].tolist() == [r.num_trailing_blank for r in want]
    lp[1, 7, :] = -np.inf                                 # a frame of -inf only: the fold cannot decide it, the exact path must
    lp[0, 8, 3] = np.inf
    want = O.ctc_greedy_search(lp)
    t, s, _, _ = h.ctc_greedy(lp)
    assert t == [r.appended for r in want] and s == [r.timestamps for r in want]
    h.set_option("ctc_one_kernel", -1)


# ---- the reference-interface mirror, fused and fine-grained, on the GPU ---------------------------------------
def test_recognizer_mirror_fused_equals_fine_grained(mid):
    m, w, _ = mid
    proj = P.OfflineProjOfB200(MID, w)
    raw = synth.make_frames(4, 30, MID.encoder_dim, 35)
    out = {}
    for fused in (True, False):
        rec = R.OfflineRecognizer(proj, fused=fused)
        ss = [rec.CreateOfflineStream() for _ in range(4)]
        for i, s in enumerate(ss):
            s.AcceptFrames(raw[i])
        rec.GetResults(ss)
        out[fused] = [(s.Tokens, s.Timestamps) for s in ss]
    want = O.greedy_search_batch(m, O.encoder_proj(m, raw), True)
    if min(r.min_gap for r in want) > NEAR_TIE:
        assert out[True] == out[False] == [(r.tokens, r.timestamps) for r in want]
    rec = R.OfflineRecognizer(proj, decodingMethod="modified_beam_search", maxActivePaths=4)
    ss = [rec.CreateOfflineStream() for _ in range(4)]
    for i, s in enumerate(ss):
        s.AcceptFrames(raw[i])
    rec.GetResults(ss)
    wb = O.modified_beam_search(m, O.encoder_proj(m, raw), 4)
    compare_streams([s.Tokens[2:] for s in ss], [s.Timestamps for s in ss], wb, "recognizer mbs")
    proj.Dispose()


# ---- full-size properties (BASELINE.json configs) ---------------------------------------------------------------
def test_full_size_cfg2_properties(built_lib):
    """cfg2 shape (B=256, T=250, K=4, V=500): determinism, shard-invariance (a stream's result does not depend
    on its batch neighbours), beam-1 == greedy, and an oracle spot check on a few streams."""
    cfg = synth.CONFIGS["cfg2"]
    m, w = model_and_weights(cfg.dims, blank_bias=cfg.blank_bias)
    h = make_handle(cfg.dims, w)
    raw = synth.make_frames(cfg.streams, cfg.frames, cfg.dims.encoder_dim, cfg.seed)
    t1, s1, sc1 = h.modified_beam_search(raw, 4)
    t2, s2, sc2 = h.modified_beam_search(raw, 4)
    assert t1 == t2 and s1 == s2 and sc1.tolist() == sc2.tolist()
    ta, sa, sca = h.modified_beam_search(np.ascontiguousarray(raw[64:128]), 4)
    assert ta == t1[64:128] and sa == s1[64:128]
    np.testing.assert_allclose(sca, sc1[64:128], atol=0)
    assert all(len(t) == len(s) and all(0 <= x < cfg.frames for x in s) and s == sorted(s) for t, s in zip(t1, s1))
    assert all(all(3 <= tok < cfg.dims.vocab_size or tok == 1 for tok in t) for t in t1)
    enc = O.encoder_proj(m, raw[:6])
    want = O.modified_beam_search(m, enc, 4)
    ex = compare_streams(t1[:6], s1[:6], want, "cfg2 spot", allow_frac=0.15)
    for b in range(6):
        if b not in ex:
            assert abs(float(sc1[b]) - want[b].score) < SCORE_TOL
    g1, _, _ = h.modified_beam_search(np.ascontiguousarray(raw[:32]), 1)
    g2, _ = h.greedy_offline(np.ascontiguousarray(raw[:32]), _native.GREEDY_PER_STREAM)
    assert sum(a != b for a, b in zip(g1, g2)) <= 1
    h.close()


def test_full_size_cfg5_ctc_properties(built_lib):
    """cfg5 per-GPU shard (B=128, T=250, V=2000): whole == concatenation of halves with carried state;
    idempotence; every emitted token non-blank and no immediate repeats within a stream's frame ids."""
    h = make_handle(SMALL, None)
    lp = synth.make_ctc_logp(128, 250, 2000, 1005, blank_bias=11.5)
    t, s, tb, pv = h.ctc_greedy(lp, trailing_blank=np.zeros(128, np.int32), prev=np.full(128, -1, np.int64))
    t2, s2, _, _ = h.ctc_greedy(lp)
    assert t == t2 and s == s2
    a = h.ctc_greedy(np.ascontiguousarray(lp[:, :125]), prev=np.full(128, -1, np.int64), trailing_blank=np.zeros(128, np.int32))
    b = h.ctc_greedy(np.ascontiguousarray(lp[:, 125:]), prev=a[3], trailing_blank=a[2], frame_offset=np.full(128, 125, np.int32))
    assert [x + y for x, y in zip(a[0], b[0])] == t and b[2].tolist() == tb.tolist()
    ids = lp.argmax(-1)
    for bi in (0, 17, 127):
        want = O.ctc_greedy_search(lp[bi:bi + 1])[0]
        assert t[bi] == want.appended and s[bi] == want.timestamps
        assert all(tok != 0 for tok in t[bi])
    assert sum(len(x) for x in t) == int(((ids != 0) & (ids != np.concatenate([np.full((128, 1), -1), ids[:, :-1]], 1))).sum())
    h.close()


def test_full_size_cfg4_properties(built_lib, monkeypatch):
    """cfg4 shape (B=256, T=250 cut to 60 here, K=4, V=5537: the persistent beam kernel - tcgen05 joiner CTAs + merge warps in one
    launch): determinism, shard-invariance, token range, an oracle spot check, ragged lengths, and - at the full 250 frames -
    bit-identical results from the per-frame launches (a different schedule of the same arithmetic)."""
    cfg = synth.CONFIGS["cfg4"]
    T = 60
    m, w = model_and_weights(cfg.dims, blank_bias=cfg.blank_bias)
    h = make_handle(cfg.dims, w, precision=_native.PREC_NAMES["bf16x3"])
    raw = synth.make_frames(cfg.streams, T, cfg.dims.encoder_dim, cfg.seed)
    t1, s1, sc1 = h.modified_beam_search(raw, 4, enc_is_raw=True)
    t2, s2, sc2 = h.modified_beam_search(raw, 4, enc_is_raw=True)
    assert t1 == t2 and s1 == s2 and sc1.tolist() == sc2.tolist()
    ta, sa, sca = h.modified_beam_search(np.ascontiguousarray(raw[128:160]), 4, enc_is_raw=True)
    assert ta == t1[128:160] and sa == s1[128:160]
    np.testing.assert_allclose(sca, sc1[128:160], atol=0)
    assert all(len(t) == len(s) and all(0 <= x < T for x in s) and s == sorted(s) for t, s in zip(t1, s1))
    assert all(all(3 <= tok < cfg.dims.vocab_size or tok == 1 for tok in t) for t in t1)
    enc = O.encoder_proj(m, raw[:4])
    want = O.modified_beam_search(m, enc, 4)
    ex = compare_streams(t1[:4], s1[:4], want, "cfg4 spot", allow_frac=0.15)
    for b in range(4):
        if b not in ex:
            assert abs(float(sc1[b]) - want[b].score) < SCORE_TOL
    lens = [T - (7 * b) % T for b in range(cfg.streams)]
    tr, sr, _ = h.modified_beam_search(raw, 4, enc_is_raw=True, lens=lens)
    assert all(all(x < lens[b] for x in sr[b]) for b in range(cfg.streams))
    assert [tr[b] for b in range(cfg.streams) if lens[b] == T] == [t1[b] for b in range(cfg.streams) if lens[b] == T]
    full = synth.make_frames(cfg.streams, cfg.frames, cfg.dims.encoder_dim, cfg.seed + 1)
    tm, sm, scm = h.modified_beam_search(full, 4, enc_is_raw=True)
    h.set_option("no_mega", 1)
    tp, sp, scp = h.modified_beam_search(full, 4, enc_is_raw=True)
    h.set_option("no_mega", 0)
    assert tm == tp and sm == sp and scm.tolist() == scp.tolist()
    assert sum(len(t) for t in tm) > cfg.streams * 10            # it decoded something
    h.close()


def test_full_size_cfg3_online_properties(built_lib):
    """cfg3 shape (512 online streams, V=2000, chunks of 8 frames, 16-CTA clusters): chunked == one long chunk, determinism,
    shard-invariance, Hyp carried through, and an oracle spot check."""
    cfg = synth.CONFIGS["cfg3"]
    m, w = model_and_weights(cfg.dims, blank_bias=cfg.blank_bias)
    h = make_handle(cfg.dims, w, precision=_native.PREC_NAMES["bf16x3"])
    B, Tc, nch = cfg.streams, cfg.frames, 4
    raw = synth.make_frames(B, Tc * nch, cfg.dims.encoder_dim, cfg.seed)
    hyp = np.zeros((B, 2), np.int64)
    toks = [[] for _ in range(B)]
    tss = [[] for _ in range(B)]
    for c in range(nch):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
        for b in range(B):
            toks[b] += t[b]
            tss[b] += [x + Tc * c for x in s[b]]
    t_all, s_all, hyp_all = h.greedy_online_chunk(raw, np.zeros((B, 2), np.int64), enc_is_raw=True)
    assert toks == t_all and tss == s_all and hyp.tolist() == hyp_all.tolist()
    t_sub, s_sub, _ = h.greedy_online_chunk(np.ascontiguousarray(raw[100:140]), np.zeros((40, 2), np.int64), enc_is_raw=True)
    assert t_sub == t_all[100:140] and s_sub == s_all[100:140]
    for b in range(B):
        tail = ([0, 0] + t_all[b])[-2:]
        assert hyp_all[b].tolist() == tail                      # ref OnlineRecognizer.cs:208: Hyp = last two tokens
        assert all(tok not in (0, 1, 2) for tok in t_all[b])    # ref :181: blank, unk and the literal 1 are never emitted
    enc = O.encoder_proj(m, raw[:6])
    res = O.greedy_search_online_chunk(m, enc, [[0, 0]] * 6, [[0, 0]] * 6)
    compare_streams(t_all[:6], s_all[:6], res, "cfg3 spot", allow_frac=0.15)
    h.close()
