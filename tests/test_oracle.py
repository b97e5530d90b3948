"""CPU: the oracle against the committed golden vectors and against the crafted quirk cases of
SURVEY.md section 3.6 (the reference has no tests of its own to inherit: section 4)."""
from pathlib import Path

import numpy as np
import pytest

from k2transducerasr_b200 import synth
from oracle import k2_oracle as O
from tests.helpers import SMALL, model_and_weights

GOLD = np.load(Path(__file__).parent / "golden" / "small_v53.npz")


def _unragged(a, n):
    return [a[i, :n[i]].tolist() for i in range(len(n))]


@pytest.fixture(scope="module")
def model():
    m, _ = model_and_weights(SMALL, blank_bias=0.6)
    return m


def test_golden_model_math(model):
    enc = O.encoder_proj(model, GOLD["raw"])
    np.testing.assert_allclose(enc, GOLD["enc"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(O.decoder(model, GOLD["y"]), GOLD["dec"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(O.joiner(model, GOLD["enc"][:, 0, :], GOLD["dec"]), GOLD["logits"], rtol=0, atol=2e-6)


def test_golden_searches(model):
    enc = GOLD["enc"]
    s = O.greedy_search_single(model, enc[0])
    assert s.tokens == GOLD["single_tokens"].tolist() and s.timestamps == GOLD["single_ts"].tolist()
    for name, res in (("compat", O.greedy_search_batch(model, enc, True)),
                      ("perstream", O.greedy_search_batch(model, enc, False)),
                      ("mbs4", O.modified_beam_search(model, enc, 4)),
                      ("mbs2", O.modified_beam_search(model, enc, 2))):
        assert [r.tokens for r in res] == _unragged(GOLD[f"{name}_tokens"], GOLD[f"{name}_ntok"]), name
        assert [r.timestamps for r in res] == _unragged(GOLD[f"{name}_ts"], GOLD[f"{name}_nts"]), name
        np.testing.assert_allclose([r.score for r in res], GOLD[f"{name}_score"], atol=1e-4)


def test_golden_ctc():
    res = O.ctc_greedy_search(GOLD["ctc_logp"], 0, frame_offset=[0, 5, 0, 0], trailing_blank=[0, 2, 0, 0])
    assert [r.appended for r in res] == _unragged(GOLD["ctc_tokens"], GOLD["ctc_ntok"])
    assert [r.timestamps for r in res] == _unragged(GOLD["ctc_ts"], GOLD["ctc_nts"])
    assert [r.num_trailing_blank for r in res] == GOLD["ctc_trailing"].tolist()
    assert res[1].appended == [] and res[1].num_trailing_blank == 22      # all blank: 2 carried + 20
    assert 5 in res[2].appended and 9 not in [t for t, ts in zip(res[2].appended, res[2].timestamps) if ts == 3]


def test_q1_transducer_argmax_ties_and_nan_go_high():
    # literal fold of ref OfflineRecognizer.cs:150-154
    def fold(row):
        tok = 0
        for k in range(1, len(row)):
            tok = tok if row[tok] > row[k] else k
        return tok
    rows = np.array([[1, 3, 3, 0], [2, 2, 2, 2], [0, np.nan, 5, 1], [9, 1, np.nan, 0], [np.nan, 0, 0, 0],
                     [1, 2, 3, np.nan]], np.float32)
    assert O.argmax_hi(rows).tolist() == [fold(r) for r in rows] == [2, 3, 2, 3, 3, 3]


def test_q2_ctc_argmax_ties_go_low():
    assert O.argmax_lo(np.array([1, 3, 3, 0], np.float32)) == 1
    assert O.argmax_lo(np.array([np.nan, np.nan], np.float32)) == 0
    assert O.argmax_lo(np.array([np.nan, -np.inf, 2, 2], np.float32)) == 2
    assert O.argmax_lo(np.array([np.nan, -np.inf, -np.inf], np.float32)) == 1


def test_q3_unk_is_not_emitted_offline_and_id1_not_online(model):
    m, w = model_and_weights(SMALL)
    w["out_b"] = w["out_b"].copy()
    w["out_b"][2] += 50.0                                    # unk always wins
    m = O.Model.from_dict(w)
    enc = synth.make_frames(2, 6, SMALL.joiner_dim, 3)
    assert all(r.appended == [] for r in O.greedy_search_batch(m, enc, True))
    w["out_b"][2] -= 50.0
    w["out_b"][1] += 50.0                                    # id 1 always wins
    m = O.Model.from_dict(w)
    off = O.greedy_search_batch(m, enc, False)
    assert all(r.appended == [1] * 6 for r in off)           # offline emits id 1 (ref :161, :268)
    on = O.greedy_search_online_chunk(m, enc, [[0, 0]] * 2, [[0, 0]] * 2)
    assert all(r.appended == [] and r.hyp == [0, 0] for r in on)   # online masks the literal 1 (ref :181)


def test_q5_q6_batch_compat_seed_and_context_flip(model):
    enc = GOLD["enc"]
    B = enc.shape[0]
    compat = O.greedy_search_batch(model, enc, True)
    per = O.greedy_search_batch(model, enc, False)
    for r in compat:
        assert r.tokens[:2 * B] == [0] * (2 * B) and r.timestamps[:2 * B] == [0] * (2 * B)   # Q5
    assert per[0].appended == O.greedy_search_single(model, enc[0]).appended
    # Q6 is observable: a stream decoded alone differs from the same stream inside a batch whenever its
    # first emission comes after a neighbour's and the [-1,0] vs [0,0] decoder outputs decide differently.
    d_init = O.decoder(model, np.array([[-1, 0]], np.int64))
    d_flip = O.decoder(model, np.array([[0, 0]], np.int64))
    assert not np.allclose(d_init, d_flip)


def test_q10_online_ctc_repeat_across_chunk_edge_is_emitted_twice():
    lp = GOLD["ctc_logp"][3:4]
    whole = O.ctc_greedy_search(lp)[0]
    a = O.ctc_greedy_search(lp[:, :10])[0]
    b = O.ctc_greedy_search(lp[:, 10:])[0]                   # prev_id reset -> 11 again
    assert whole.appended.count(11) + 1 == (a.appended + b.appended).count(11)
    b2 = O.ctc_greedy_search(lp[:, 10:], prev=a.hyp)[0]      # carried prev (our option) = whole utterance
    assert a.appended + b2.appended == whole.appended


def test_beam_merge_cases():
    """(v) of SURVEY 8c: `A+blank` vs `A+unk` collapse, and `prefix+token == other hyp` log-adds."""
    m, w = model_and_weights(SMALL)
    w["out_b"] = w["out_b"].copy()
    w["out_b"][0] += 6.0
    w["out_b"][2] += 6.0                                     # blank and unk dominate: both keep ys unchanged
    m = O.Model.from_dict(w)
    enc = synth.make_frames(1, 1, SMALL.joiner_dim, 8)
    r = O.modified_beam_search(m, enc, 2)[0]
    lp = O.log_softmax(O.joiner(m, enc[:, 0], O.decoder(m, None, 1)))[0]
    assert r.appended == [] and abs(r.score - float(O.logaddexp32(lp[0], lp[2]))) < 1e-6
    # beam=1 equals greedy (tie rule consistent with Q1)
    m2, _ = model_and_weights(SMALL, blank_bias=0.6)
    enc2 = GOLD["enc"]
    assert [x.appended for x in O.modified_beam_search(m2, enc2, 1)] == \
           [x.appended for x in O.greedy_search_batch(m2, enc2, False)]


def test_bf16_emulation_modes_are_close(model):
    y = GOLD["y"]
    ref = O.decoder(model, y)
    m3 = O.Model.from_dict(synth.make_weights(SMALL, blank_bias=0.6), prec="bf16x3")
    m1 = O.Model.from_dict(synth.make_weights(SMALL, blank_bias=0.6), prec="bf16")
    assert np.abs(O.decoder(m3, y) - ref).max() < 2e-5
    assert 1e-5 < np.abs(O.decoder(m1, y) - ref).max() < 5e-2
    assert O.round_bf16(np.array([1.00390625], np.float32))[0] == np.float32(1.0)    # RNE tie -> even


def test_stack_unstack_states_restate_the_reference_loops():
    """ref OnlineProjOfZipformer2.cs:236-246 (stack: (x*B + n)*A + a  <-  n-th item [x*A + a]) and :399-404 (unstack); batch on
    axis 1 of [X, B, A], so stacking is a transpose of the two leading axes and unstacking inverts it for the same A."""
    rng = np.random.default_rng(0)
    B, shapes = 3, [(4, 6), (1, 5), (7, 1)]                      # (X, A) per cache tensor
    items = [[rng.standard_normal(x * a).astype(np.float32) for x, a in shapes] for _ in range(B)]
    axis = [a for _, a in shapes]
    st = O.stack_states(items, axis)
    for i, (x, a) in enumerate(shapes):
        want = np.stack([items[n][i].reshape(x, a) for n in range(B)], axis=1).reshape(-1)
        np.testing.assert_array_equal(st[i], want)
    back = O.unstack_states(st, B, axis)
    for n in range(B):
        for i in range(len(shapes)):
            np.testing.assert_array_equal(back[n][i], items[n][i])
    # a different A in the two directions (the reference's cached_nonlin_attn, :250 vs :409) is NOT an inverse pair for B > 1
    st2 = O.stack_states(items, [2, 5, 1])
    back2 = O.unstack_states(st2, B, [6, 5, 1])
    assert not all(np.array_equal(back2[n][0], items[n][0]) for n in range(B))


def test_ragged_wrapper_is_per_stream_decoding():
    m, w = model_and_weights(SMALL, blank_bias=0.6)
    enc = synth.make_frames(4, 9, m.J, 5)
    lens = [9, 0, 4, 7]
    got = O.ragged(O.modified_beam_search, m, enc, lens, 3)
    for b, n in enumerate(lens):
        one = O.modified_beam_search(m, enc[b:b + 1, :n], 3)[0]
        assert got[b].appended == one.appended and got[b].timestamps == one.timestamps
        assert all(t < n for t in got[b].timestamps)
    assert got[1].appended == []
