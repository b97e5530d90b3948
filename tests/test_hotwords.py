"""CPU: contextual biasing (hot words) - the dense automaton of k2transducerasr_b200/hotwords.py against the definition restated
by brute force, and the oracle's biased beam search against the textbook search with the brute-force boost (no automaton)."""
import numpy as np
import pytest

from k2transducerasr_b200 import synth
from k2transducerasr_b200.hotwords import ContextGraph, brute_force_boost
from oracle import k2_oracle as O
from oracle import naive_torch as N
from tests.helpers import SMALL, model_and_weights


def test_automaton_equals_brute_force_definition():
    rng = np.random.default_rng(0)
    words = [[3, 4, 5], [3, 4], [4, 5, 6, 7], [9], [5, 5, 5], [4, 4, 3], []]
    g = ContextGraph.build(words, 12, 2.0)
    assert g.next.shape == g.delta.shape == (g.num_states, 12) and g.residual[0] == 0
    for _ in range(3000):
        toks = rng.integers(2, 10, size=rng.integers(0, 14)).tolist()
        assert abs(g.total_boost(toks) - brute_force_boost(words, toks, 2.0)) < 1e-5, toks
    # a match that breaks gives the boost back; a completed hot word keeps it
    assert g.total_boost([4, 5, 6], finalize=False) == pytest.approx(6.0) and g.total_boost([4, 5, 6]) == 0.0
    assert g.total_boost([4, 5, 6, 2]) == 0.0 and g.total_boost([4, 5, 6, 7, 2]) == pytest.approx(8.0)
    assert g.total_boost([3, 4, 5]) == pytest.approx(4.0)          # [3, 4] completes first; the 5 starts [5, 5, 5] and is revoked
    with pytest.raises(ValueError):
        ContextGraph.build([[99]], 12)


@pytest.mark.parametrize("beam", [2, 4])
def test_biased_beam_search_equals_textbook_restatement(beam):
    m, w = model_and_weights(SMALL, blank_bias=0.6)
    tm = N.TorchModel(w)
    enc = O.encoder_proj(m, synth.make_frames(6, 30, SMALL.encoder_dim, 11))
    plain = O.modified_beam_search(m, enc, beam)
    # hot words made of tokens the plain search emits (so the boost has something to act on), plus one that never occurs
    seen = [t for r in plain for t in r.appended]
    words = [seen[0:2], seen[3:6], seen[7:8], [50, 51, 52]]
    g = ContextGraph.build(words, SMALL.vocab_size, 2.5)
    want = O.modified_beam_search(m, enc, beam, context_graph=g)
    ndiff = 0
    for b, r in enumerate(want):
        toks, ts, lp = N.beam_search(tm, enc[b], beam, hotwords=words, hot_score=2.5)
        if toks == r.appended and ts == r.timestamps:
            assert abs(lp - r.score) < 1e-4
        else:
            assert r.min_gap < 1e-5, (b, r.min_gap)
            ndiff += 1
    assert ndiff <= 1
    assert any(a.appended != b.appended or abs(a.score - b.score) > 1.0 for a, b in zip(plain, want)), "the boost changed nothing"
    # score 0 is the unbiased search; chunked == whole with the automaton state carried
    g0 = ContextGraph.build(words, SMALL.vocab_size, 0.0)
    z = O.modified_beam_search(m, enc, beam, context_graph=g0)
    assert [r.appended for r in z] == [r.appended for r in plain] and [r.score for r in z] == [r.score for r in plain]
    st, off = None, 0
    for n in (7, 1, 22):
        res, st = O.modified_beam_search(m, enc[:, off:off + n], beam, init=st, frame_offset=[off] * 6, context_graph=g, return_state=True)
        off += n
    assert [r.appended for r in res] == [r.appended for r in want] and [r.score for r in res] == [r.score for r in want]
