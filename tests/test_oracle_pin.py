"""Pins oracle/k2_oracle.py (CPU, no GPU needed).

The reference cannot be executed in this image and ships no golden vectors, so the numpy oracle is
pinned against a SECOND implementation that shares no code with it (oracle/naive_torch.py): the model math against the
PyTorch operators the ONNX graphs are exported from, the searches against textbook restatements (full sort, dict keyed by the
token tuple, literal scalar loops of the reference's comparison expressions) - on random inputs and on crafted exact ties.
Also here: chunked == whole-utterance beam search, max_sym_per_frame, and the parity checker's own rules.
"""
import numpy as np
import pytest
import torch

from k2transducerasr_b200 import synth
from oracle import k2_oracle as O
from oracle import naive_torch as N
from oracle import parity as P
from tests.helpers import MID, SMALL, model_and_weights

torch.set_num_threads(4)


def _models(dims, blank_bias=0.0, seed=synth.WEIGHT_SEED, **kw):
    m, w = model_and_weights(dims, blank_bias=blank_bias, seed=seed, **kw)
    return m, N.TorchModel(w, blank_id=dims.blank_id, unk_id=dims.unk_id, context_size=dims.context_size,
                           neg_id_wrap=kw.get("neg_id_wrap", False)), w


@pytest.mark.parametrize("dims", [SMALL, MID])
@pytest.mark.parametrize("wrap", [False, True])
def test_model_math_equals_torch_functional(dims, wrap):
    """decoder (embedding * mask -> conv1d(groups = D/4) -> relu -> linear), joiner (linear(tanh(a + b))), encoder_proj and
    log_softmax of the oracle against torch.nn.functional: 1e-5 absolute (accumulation order is all that differs)."""
    m, tm, w = _models(dims, neg_id_wrap=wrap)
    rng = np.random.default_rng(5)
    V = dims.vocab_size
    y = rng.integers(0, V, size=(64, 2))
    y[:8, 0] = -1                                  # the offline seed {-1, blank} (Q4)
    y[8, :] = [V - 1, V - 1]
    d_o = O.decoder(m, y)
    d_t = tm.decoder(y).numpy()
    assert d_o.shape == d_t.shape
    assert np.abs(d_o - d_t).max() < 1e-5
    assert np.abs(O.decoder(m, None, 3) - tm.decoder([[-1, dims.blank_id]] * 3).numpy()).max() < 1e-5
    enc = rng.standard_normal((64, dims.joiner_dim)).astype(np.float32)
    lg_o = O.joiner(m, enc, d_o)
    lg_t = tm.joiner(enc, torch.from_numpy(d_o)).numpy()
    assert np.abs(lg_o - lg_t).max() < 1e-5
    raw = rng.standard_normal((3, 7, dims.encoder_dim)).astype(np.float32)
    assert np.abs(O.encoder_proj(m, raw) - tm.encoder_proj(raw).numpy()).max() < 1e-5
    assert np.abs(O.log_softmax(lg_o) - torch.log_softmax(torch.from_numpy(lg_o), -1).numpy()).max() < 2e-6


def test_negative_id_semantics():
    """mask mode: a negative id contributes a zero embedding row; wrap mode: ONNX Gather semantics, -1 addresses row V-1."""
    m, tm, w = _models(SMALL)
    z = O.decoder_conv(m, np.array([[-1, 3]]))
    w0 = dict(w); w0["emb"] = w["emb"].copy()
    m2 = O.Model.from_dict(w0)
    assert np.array_equal(z, O.decoder_conv(m2, np.array([[-1, 3]])))
    mw, tw, _ = _models(SMALL, neg_id_wrap=True)
    V = SMALL.vocab_size
    assert np.allclose(O.decoder(mw, np.array([[-1, 3]])), O.decoder(mw, np.array([[V - 1, 3]])))
    assert not np.allclose(O.decoder(m, np.array([[-1, 3]])), O.decoder(m, np.array([[V - 1, 3]])))


@pytest.mark.parametrize("beam", [1, 2, 4, 8])
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_beam_search_equals_textbook_restatement(beam, seed):
    """np.partition / lexsort / index-list search of the oracle == full sort + dict-of-tuples search on torch math, stream by
    stream; a stream may differ only where the oracle itself met a margin below 1e-5 (the two arithmetic back-ends differ by
    a few ulp)."""
    m, tm, _ = _models(SMALL, blank_bias=0.6)
    enc = O.encoder_proj(m, synth.make_frames(6, 30, SMALL.encoder_dim, seed))
    want = O.modified_beam_search(m, enc, beam)
    ndiff = 0
    for b, r in enumerate(want):
        toks, ts, lp = N.beam_search(tm, enc[b], beam)
        if toks == r.appended and ts == r.timestamps:
            assert abs(lp - r.score) < 1e-4
        else:
            assert r.min_gap < 1e-5, (b, r.min_gap, toks, r.appended)
            ndiff += 1
    assert ndiff <= 1


def _tie_model(V=9, J=16, D=16):
    """All-zero joiner weight: logits == out_b exactly in every implementation, so crafted ties are exact."""
    dims = synth.ModelDims(vocab_size=V, joiner_dim=J, decoder_dim=D, encoder_dim=0)
    w = synth.make_weights(dims)
    w["out_w"] = np.zeros_like(w["out_w"])
    return dims, w


def test_crafted_ties_beam_and_greedy():
    """Exact ties: the transducer argmax goes to the LARGER index (Q1), beam search ranks value-desc then flat-index-desc, merges
    log-add; both implementations must agree token for token, score to the last bit of a float32 log-add."""
    dims, w = _tie_model()
    w["out_b"] = np.array([0.5, -3, -3, 0.5, 0.25, 0.5, -2, 0.25, -1], np.float32)   # three-way tie 0 / 3 / 5, two-way 4 / 7
    m = O.Model.from_dict(w)
    tm = N.TorchModel(w)
    enc = synth.make_frames(2, 6, dims.joiner_dim, 3)
    g = O.greedy_search_single(m, enc[0])
    assert g.appended == [5] * 6 and N.greedy_single(tm, enc[0])[0] == [5] * 6
    for beam in (1, 2, 3, 4, 8):
        want = O.modified_beam_search(m, enc, beam)
        for b, r in enumerate(want):
            toks, ts, lp = N.beam_search(tm, enc[b], beam)
            assert toks == r.appended and ts == r.timestamps, (beam, b, toks, r.appended)
            assert abs(lp - r.score) < 2e-5
    # blank / unk extensions merge with their parent (A+blank == A+unk): scores are log-added
    w["out_b"] = np.array([1.0, -9, 1.0, -9, 0.0, -9, -9, -9, -9], np.float32)
    m, tm = O.Model.from_dict(w), N.TorchModel(w)
    r = O.modified_beam_search(m, enc[:1], 4)[0]
    toks, ts, lp = N.beam_search(tm, enc[0], 4)
    assert toks == r.appended and abs(lp - r.score) < 2e-5
    assert all(len(rec) <= 4 for rec in r.history)


@pytest.mark.parametrize("msf", [1, 2, 3])
def test_greedy_loops_equal_literal_restatement(msf):
    """Single-stream greedy incl. max_sym_per_frame (ref OfflineRecognizer.cs:127-179), one stream of the online loop (mask
    {blank, unk, 1}, Hyp seed), and CTC greedy, against the literal scalar loops."""
    m, tm, _ = _models(SMALL, blank_bias=0.3)
    enc = O.encoder_proj(m, synth.make_frames(4, 25, SMALL.encoder_dim, 77))
    for b in range(4):
        r = O.greedy_search_single(m, enc[b], max_sym_per_frame=msf)
        toks, ts = N.greedy_single(tm, enc[b], max_sym_per_frame=msf)
        if toks != r.appended or ts != r.timestamps:
            assert r.min_gap < 1e-5
        if msf > 1:
            assert max(np.bincount(np.asarray(r.timestamps, np.int64), minlength=1)) <= msf
    if msf == 1:
        res = O.greedy_search_online_chunk(m, enc, [[0, 0]] * 4, [[0, 0]] * 4)
        for b, r in enumerate(res):
            toks, ts = N.greedy_single(tm, enc[b], extra_mask=1, hyp=[0, 0])
            assert (toks == r.appended and ts == r.timestamps) or r.min_gap < 1e-5
        lp = synth.make_ctc_logp(3, 40, 31, 5, blank_bias=2.0)
        lp[1, 7, :] = -3.0
        lp[1, 7, 4] = lp[1, 7, 9] = -0.5              # exact tie -> lowest index (Q2)
        for b, r in enumerate(O.ctc_greedy_search(lp)):
            toks, ts, tb = N.ctc_greedy(lp[b])
            assert toks == r.appended and ts == r.timestamps and tb == r.num_trailing_blank


@pytest.mark.parametrize("beam", [2, 4])
def test_streaming_beam_search_equals_whole_utterance(beam):
    """Hypotheses carried from chunk to chunk (what k2b_modified_beam_search_online_chunk does on the device): decoding 5 chunks of
    uneven length equals decoding the utterance whole - tokens, absolute timestamps, scores, beam history."""
    m, _, _ = _models(SMALL, blank_bias=0.6)
    enc = O.encoder_proj(m, synth.make_frames(5, 37, SMALL.encoder_dim, 21))
    seed = O.mbs_seed(m, 5, [[0, 0]] * 5)
    whole = O.modified_beam_search(m, enc, beam, init=seed, extra_mask=1)
    state, off = seed, 0
    hist = [[] for _ in range(5)]
    for n in (8, 8, 1, 13, 7):
        res, state = O.modified_beam_search(m, enc[:, off:off + n], beam, init=state, frame_offset=[off] * 5, extra_mask=1,
                                            return_state=True)
        for b in range(5):
            hist[b] += res[b].history
        off += n
    for b in range(5):
        assert res[b].tokens == whole[b].tokens and res[b].timestamps == whole[b].timestamps
        assert res[b].score == whole[b].score and hist[b] == whole[b].history
        assert res[b].hyp == whole[b].tokens[-2:]


def test_parity_checker_rules():
    """oracle/parity.py: identical -> 100 %; a difference is excused only by the margin AT the first divergent frame."""
    m, _, _ = _models(SMALL, blank_bias=0.6)
    enc = O.encoder_proj(m, synth.make_frames(3, 20, SMALL.encoder_dim, 9))
    want = O.greedy_search_batch(m, enc, compat=False)
    toks = [list(r.appended) for r in want]
    ts = [list(r.timestamps) for r in want]
    rep = P.compare(toks, ts, want, 20)
    assert rep.streams_identical == 3 and rep.frames_identical_pct == 100.0 and not rep.near_tie_frames and not rep.unexplained
    # drop the first symbol of stream 1: the divergence is at its timestamp, whose margin is no near tie -> unexplained
    bad_t, bad_s = [list(t) for t in toks], [list(s) for s in ts]
    t_first = bad_s[1][0]
    del bad_t[1][0], bad_s[1][0]
    rep = P.compare(bad_t, bad_s, want, 20)
    assert rep.unexplained and rep.unexplained[0][:2] == (1, t_first) and rep.streams_identical == 2
    assert rep.frames_identical == 59
    with pytest.raises(AssertionError):
        rep.assert_ok("crafted")
    # the same difference IS excused when the oracle's margin at that very frame is a near tie - and only then
    want[1].frame_gap[t_first] = 5e-5
    rep = P.compare(bad_t, bad_s, want, 20)
    assert not rep.unexplained and rep.near_tie_frames[0][:2] == (1, t_first)
    want[1].frame_gap[t_first] = 1.0
    want[1].frame_gap[(t_first + 1) % 20] = 1e-6         # a near tie elsewhere in the stream excuses nothing
    assert P.compare(bad_t, bad_s, want, 20).unexplained


def test_parity_checker_beam_history():
    """Beam search is judged on the beam: the back-pointer rows must equal the oracle's history; a swapped pair of slots at a
    frame is located at that frame."""
    m, _, _ = _models(SMALL, blank_bias=0.6)
    enc = O.encoder_proj(m, synth.make_frames(2, 12, SMALL.encoder_dim, 19))
    want = O.modified_beam_search(m, enc, 4)
    bp = np.zeros((2, 12, 4), np.int32)
    for b, r in enumerate(want):
        for t, rec in enumerate(r.history):
            for k, (par, tok) in enumerate(rec):
                bp[b, t, k] = (par << 28) | (tok + 1)
    toks, ts = [r.appended for r in want], [r.timestamps for r in want]
    rep = P.compare(toks, ts, want, 12, got_score=[r.score for r in want], bp=bp)
    assert rep.beam_frames_identical == 24 and not rep.unexplained and not rep.near_tie_frames
    t_sw = next(t for t, rec in enumerate(want[0].history) if len(rec) >= 2 and rec[0] != rec[1])
    bp2 = bp.copy()
    bp2[0, t_sw, [0, 1]] = bp[0, t_sw, [1, 0]]
    rep = P.compare(toks, ts, want, 12, bp=bp2)
    assert (rep.unexplained + rep.near_tie_frames)[0][:2] == (0, t_sw)
    assert rep.beam_frames_identical == 12 + t_sw
