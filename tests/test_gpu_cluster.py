"""GPU (-m gpu): the persistent cluster kernel (tcgen05 precisions) of modified_beam_search against the oracle.

bf16x3 (split-bf16, three MMAs, fp32 accumulate) is the parity-grade tensor-core mode: token sequences must
match the fp32 oracle except on near ties, scores within 1e-3. bf16 is the fast mode: it is compared against
the oracle restated with bf16-rounded joiner operands (same tie rule), and its distance from fp32 is reported
with a stated tolerance."""
import numpy as np
import pytest

from k2transducerasr_b200 import _native, synth
from oracle import k2_oracle as O
from tests.helpers import MID, SCORE_TOL, compare_streams, model_and_weights

pytestmark = pytest.mark.gpu

BF16_SCORE_TOL = 0.25     # stated bf16 tolerance on the un-normalised hypothesis log-prob after 40 frames


def make(dims, w, prec):
    h = _native.Handle(vocab_size=dims.vocab_size, joiner_dim=dims.joiner_dim, decoder_dim=dims.decoder_dim,
                       encoder_dim=dims.encoder_dim, precision=_native.PREC_NAMES[prec])
    h.load_weights(w)
    return h


@pytest.fixture(scope="module")
def setup(built_lib):
    m, w = model_and_weights(MID, blank_bias=0.99)
    raw = synth.make_frames(19, 40, MID.encoder_dim, 34)
    enc = O.encoder_proj(m, raw)
    return m, w, raw, enc


@pytest.mark.parametrize("beam", [4, 1, 2, 8, 3])
def test_cluster_bf16x3_matches_fp32_oracle(setup, beam):
    m, w, raw, enc = setup
    h = make(MID, w, "bf16x3")
    want = O.modified_beam_search(m, enc, beam)
    t, s, sc = h.modified_beam_search(raw, beam)
    bp = h.debug_backpointers(raw.shape[0], raw.shape[1], beam)     # the beam history: divergences are located frame by frame
    ex = compare_streams(t, s, want, f"cluster bf16x3 beam={beam}", allow_frac=0.12, bp=bp, scores=sc)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
    # projected frames in (no encoder_proj on device) give the same answer
    t2, s2, sc2 = h.modified_beam_search(enc, beam, enc_is_raw=False)
    assert sum(a != b for a, b in zip(t, t2)) <= len(ex) + 1
    h.close()


def test_cluster_bf16_matches_bf16_oracle_and_is_near_fp32(setup):
    m, w, raw, enc = setup
    h = make(MID, w, "bf16")
    mb = O.Model.from_dict(w, prec_joiner="bf16", prec_enc="bf16")
    want = O.modified_beam_search(mb, O.encoder_proj(mb, raw), 4)
    t, s, sc = h.modified_beam_search(raw, 4)
    ex = compare_streams(t, s, want, "cluster bf16 vs bf16 oracle", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < 5e-3
    ref = O.modified_beam_search(m, enc, 4)
    same = sum(a == r.appended for a, r in zip(t, ref))
    print(f"bf16 vs fp32 oracle: {same}/{len(ref)} streams identical; max |score diff| "
          f"{max(abs(float(a) - r.score) for a, r in zip(sc, ref)):.4f}")
    assert max(abs(float(a) - r.score) for a, r in zip(sc, ref)) < BF16_SCORE_TOL
    h.close()


def test_cluster_equals_per_step_path(setup):
    """Same library, two engines: per-step fp32 CUDA-core kernels vs the persistent tcgen05 cluster kernel."""
    m, w, raw, enc = setup
    hf = make(MID, w, "fp32")
    hx = make(MID, w, "bf16x3")
    tf, sf, scf = hf.modified_beam_search(raw, 4)
    tx, sx, scx = hx.modified_beam_search(raw, 4)
    diff = [b for b in range(len(tf)) if tf[b] != tx[b] or sf[b] != sx[b]]
    want = O.modified_beam_search(m, enc, 4)
    assert all(want[b].min_gap < 1e-4 for b in diff), diff
    np.testing.assert_allclose(np.delete(scf, diff), np.delete(scx, diff), atol=SCORE_TOL)
    hf.close(); hx.close()


def test_cluster_ragged_stream_count_and_switching(setup):
    """B not a multiple of the streams-per-cluster; precision switched on a live handle."""
    m, w, raw, enc = setup
    h = make(MID, w, "fp32")
    t0, s0, _ = h.modified_beam_search(raw[:5], 4)
    h.set_precision("bf16x3")
    t1, s1, _ = h.modified_beam_search(raw[:5], 4)
    want = O.modified_beam_search(m, enc[:5], 4)
    compare_streams(t1, s1, want, "cluster B=5", allow_frac=0.15)
    t2, s2, _ = h.modified_beam_search(raw[:1, :3], 4)
    compare_streams(t2, s2, O.modified_beam_search(m, enc[:1, :3], 4), "cluster B=1 T=3", allow_frac=0.15)
    h.close()


def test_cluster_full_size_cfg2(built_lib):
    cfg = synth.CONFIGS["cfg2"]
    m, w = model_and_weights(cfg.dims, blank_bias=cfg.blank_bias)
    h = make(cfg.dims, w, "bf16x3")
    raw = synth.make_frames(cfg.streams, cfg.frames, cfg.dims.encoder_dim, cfg.seed)
    t1, s1, sc1 = h.modified_beam_search(raw, 4)
    t2, s2, sc2 = h.modified_beam_search(raw, 4)
    assert t1 == t2 and s1 == s2 and sc1.tolist() == sc2.tolist()          # deterministic
    ta, sa, sca = h.modified_beam_search(np.ascontiguousarray(raw[64:128]), 4)
    assert ta == t1[64:128] and sa == s1[64:128]                            # independent of batch neighbours
    enc = O.encoder_proj(m, raw[:8])
    want = O.modified_beam_search(m, enc, 4)
    ex = compare_streams(t1[:8], s1[:8], want, "cluster cfg2 spot", allow_frac=0.15)
    for b in range(8):
        if b not in ex:
            assert abs(float(sc1[b]) - want[b].score) < SCORE_TOL
    h.close()


@pytest.mark.parametrize("n", [1, 127, 128, 700])
def test_encoder_proj_tcgen05(setup, n):
    """TMA-fed tcgen05 encoder_proj: split-bf16 x3 is fp32-grade; bf16 matches the bf16-restated oracle."""
    m, w, raw, enc = setup
    x = synth.make_frames(1, n, MID.encoder_dim, 77 + n)[0]
    want = O.encoder_proj(m, x)
    h = make(MID, w, "bf16x3")
    np.testing.assert_allclose(h.encoder_proj(x), want, rtol=0, atol=3e-5)
    h.set_precision("bf16")
    mb = O.Model.from_dict(w, prec_enc="bf16")
    got = h.encoder_proj(x)
    np.testing.assert_allclose(got, O.encoder_proj(mb, x), rtol=0, atol=3e-5)
    assert np.abs(got - want).max() > 1e-4
    h.close()


def test_cluster_greedy_single_per_stream_and_online(setup):
    """greedy_search on the cluster kernel (beam 1, same tie rule): offline single / per-stream and online chunks."""
    m, w, raw, enc = setup
    h = make(MID, w, "bf16x3")
    h.greedy_offline(enc[:1, :2], _native.GREEDY_SINGLE)   # first tensor-path call builds the tables (one-off launches)
    n_before = h.launch_count()
    t, s = h.greedy_offline(enc[:1], _native.GREEDY_SINGLE)
    assert h.launch_count() - n_before <= 4            # exp2x + cluster kernel + back-trace, not 3 launches per frame
    compare_streams(t, s, [O.greedy_search_single(m, enc[0])], "cluster greedy single", allow_frac=0.15)
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM)
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=False), "cluster greedy per_stream", allow_frac=0.12)
    # BATCH_COMPAT keeps the reference's whole-batch coupling (Q6) on the per-frame path
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT)
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=True), "compat on per-frame path", allow_frac=0.12)
    # online chunks: Hyp carried between calls, id 1 masked, chunk-local timestamps
    B, Tc = raw.shape[0], 8
    hyp = np.zeros((B, 2), np.int64)
    ohyp, otoks = [[0, 0]] * B, [[0, 0]] * B
    for c in range(4):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp)
        res = O.greedy_search_online_chunk(m, enc[:, Tc * c:Tc * c + Tc], ohyp, otoks)
        ex = compare_streams(t, s, res, f"cluster online chunk {c}", allow_frac=0.12)
        if ex:
            break
        ohyp, otoks = [r.hyp for r in res], [r.tokens for r in res]
        assert hyp.tolist() == ohyp
    h.close()


def test_cluster_online_masks_id_1(setup):
    m, w, raw, enc = setup
    w2 = {k: (None if v is None else v.copy()) for k, v in w.items()}
    w2["out_b"][1] += 50.0                                   # id 1 wins every frame
    h = make(MID, w2, "bf16x3")
    t, s = h.greedy_offline(enc[:3, :6], _native.GREEDY_PER_STREAM)
    assert t == [[1] * 6] * 3                                # offline emits id 1 (ref OfflineRecognizer.cs:161)
    t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(enc[:3, :6]), np.zeros((3, 2), np.int64), enc_is_raw=False)
    assert t == [[]] * 3 and hyp.tolist() == [[0, 0]] * 3    # online masks the literal 1 (ref OnlineRecognizer.cs:181)
    h.close()


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_cta_pair_variant_matches_oracle(setup, monkeypatch, prec):
    """K2B_PAIR=1: the CTAs 2i, 2i+1 of a cluster issue tcgen05.mma.cta_group::2 (M = 256, the B operand split between them, so
    each builds 16 of the 32 hypothesis rows). Opt-in (measured slower); must decode like the default kernel."""
    m, w, raw, enc = setup
    h = make(MID, w, prec)
    h.set_option("pair", 1)
    if prec == "bf16x3":
        mo, want_enc = m, enc
    else:
        mo = O.Model.from_dict(w, prec_joiner="bf16", prec_enc="bf16")
        want_enc = O.encoder_proj(mo, raw)
    for beam in (4, 1, 8):
        want = O.modified_beam_search(mo, want_enc, beam)
        t, s, sc = h.modified_beam_search(raw, beam)
        ex = compare_streams(t, s, want, f"pair {prec} beam={beam}", allow_frac=0.15)
        for b, r in enumerate(want):
            if b not in ex:
                assert abs(float(sc[b]) - r.score) < (SCORE_TOL if prec == "bf16x3" else 5e-3)
    h.close()


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_w_hi_in_tensor_memory_is_bit_identical(setup, prec):
    """J = 512: six (split-bf16 x3) / all eight (bf16) k-blocks of W_hi are read from tensor memory by TS MMAs in a fully unrolled
    issue loop; k2b_set_option("wh_tmem_kb", 0) selects the general kernel (everything in shared memory, SS MMAs). Same products,
    same accumulation order: tokens, timestamps and scores must be identical bit for bit, for every beam and for greedy."""
    m, w, raw, enc = setup
    h = make(MID, w, prec)
    for beam in (4, 1, 2, 8):
        h.set_option("wh_tmem_kb", -1)
        t1, s1, sc1 = h.modified_beam_search(raw, beam)
        g1, gs1 = h.greedy_offline(enc, _native.GREEDY_PER_STREAM, enc_is_raw=False)
        h.set_option("wh_tmem_kb", 0)
        t0, s0, sc0 = h.modified_beam_search(raw, beam)
        g0, gs0 = h.greedy_offline(enc, _native.GREEDY_PER_STREAM, enc_is_raw=False)
        assert t1 == t0 and s1 == s0 and np.array_equal(np.asarray(sc1), np.asarray(sc0)), f"{prec} beam {beam}"
        assert g1 == g0 and gs1 == gs0
    h.close()


def test_single_greedy_register_kernel(setup):
    """Up to 16 streams of at least 16 frames (cfg1 is one stream): one 8-CTA cluster per stream, the joiner weight in registers,
    fp32 FMA (single_greedy.cu). Against the oracle in SINGLE / PER_STREAM / BATCH_COMPAT mode (the latter: two passes, the second
    from a data-dependent frame), with ragged lengths, as online chunks of 16 frames with Hyp carried and the literal-1 mask, with
    an exact tie (larger index wins) - and the same calls with the kernel switched off must decode the same symbols."""
    m, w, raw, enc = setup
    h = make(MID, w, "bf16x3")
    for B in (1, 3, 16):
        e = np.ascontiguousarray(enc[:B])
        h.greedy_offline(e, _native.GREEDY_PER_STREAM, enc_is_raw=False)      # (first call: the memoised decoder, the packed weights)
        n0 = h.launch_count()
        t, s = h.greedy_offline(e, _native.GREEDY_PER_STREAM, enc_is_raw=False)
        n_on = h.launch_count() - n0
        want = O.greedy_search_batch(m, e, compat=False)
        compare_streams(t, s, want, f"single greedy per-stream B={B}", allow_frac=0.12, T=e.shape[1])
        h.set_option("single_greedy", 0)
        n0 = h.launch_count()
        t0, s0 = h.greedy_offline(e, _native.GREEDY_PER_STREAM, enc_is_raw=False)
        assert n_on < h.launch_count() - n0, "the register kernel needs no back-trace launch: it must have been the engine"
        h.set_option("single_greedy", 1)
        compare_streams(t0, s0, want, f"cluster greedy per-stream B={B}", allow_frac=0.12, T=e.shape[1])
        tc, sc = h.greedy_offline(e, _native.GREEDY_BATCH_COMPAT, enc_is_raw=False)
        compare_streams(tc, sc, O.greedy_search_batch(m, e, compat=True), f"single greedy batch-compat B={B}", allow_frac=0.12, T=e.shape[1])
    t, s = h.greedy_offline(raw[:1], _native.GREEDY_SINGLE)              # raw frames: encoder_proj in front
    compare_streams(t, s, [O.greedy_search_single(m, enc[0])], "single greedy SINGLE", allow_frac=1.0, T=enc.shape[1])
    lens = [40, 17, 0, 33, 25]
    t, s = h.greedy_offline(np.ascontiguousarray(enc[:5]), _native.GREEDY_PER_STREAM, enc_is_raw=False, lens=lens)
    want = [O.greedy_search_single(m, enc[b, :lens[b]]) for b in range(5)]
    compare_streams(t, s, want, "single greedy ragged", allow_frac=0.2, T=enc.shape[1])
    hyp = np.zeros((4, 2), np.int64)
    ohyp, otoks = [[0, 0]] * 4, [[0, 0] for _ in range(4)]
    for c in range(2):                                                    # online chunks of 16 frames
        ch = np.ascontiguousarray(enc[:4, 16 * c:16 * c + 16])
        t, s, hyp = h.greedy_online_chunk(ch, hyp, enc_is_raw=False)
        res = O.greedy_search_online_chunk(m, ch, ohyp, otoks)
        if compare_streams(t, s, res, f"single greedy online chunk {c}", allow_frac=0.25):
            break
        ohyp, otoks = [r.hyp for r in res], [r.tokens for r in res]
        assert hyp.tolist() == ohyp
    h.close()
    w2 = {k: (None if v is None else v.copy()) for k, v in w.items()}
    w2["out_w"][:] = 0.0                                                  # logits == bias in every implementation: exact ties
    w2["out_b"][:] = -3.0
    w2["out_b"][[7, 130, 499]] = 2.5
    h = make(MID, w2, "bf16x3")
    t, s = h.greedy_offline(np.ascontiguousarray(enc[:2, :20]), _native.GREEDY_PER_STREAM, enc_is_raw=False)
    assert t == [[499] * 20] * 2 and s == [list(range(20))] * 2
    h.close()


@pytest.mark.parametrize("prec", ["bf16x3", "fp32"])
def test_ragged_lengths_freeze_streams(setup, prec):
    """k2b_set_encoder_out_lens (the seam's encoder_out_lens, which the reference never consumes): stream b is decoded over its
    first lens[b] frames only - both engines (persistent cluster kernel, per-frame kernels), beam search and per-stream greedy,
    lengths 0 and T included; the setting is consumed by one call."""
    m, w, raw, enc = setup
    h = make(MID, w, prec)
    B, T = 11, 24
    lens = [24, 0, 5, 17, 1, 24, 9, 13, 2, 20, 7]
    want = O.ragged(O.modified_beam_search, m, enc[:B, :T], lens, 4)
    t, s, sc = h.modified_beam_search(raw[:B, :T], 4, lens=lens)
    ex = compare_streams(t, s, want, f"ragged mbs {prec}", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
        assert all(x < lens[b] for x in s[b])
    t2, s2, _ = h.modified_beam_search(raw[:B, :T], 4)               # consumed: the next call decodes all T frames
    full = O.modified_beam_search(m, enc[:B, :T], 4)
    compare_streams(t2, s2, full, f"after ragged {prec}", allow_frac=0.15)
    wantg = O.ragged(O.greedy_search_batch, m, enc[:B, :T], lens, compat=False)
    t, s = h.greedy_offline(raw[:B, :T], _native.GREEDY_PER_STREAM, lens=lens)
    compare_streams(t, s, wantg, f"ragged greedy {prec}", allow_frac=0.15)
    if prec == "bf16x3":            # T >= 32 raw frames: the time-chunk pipelined host call (lengths cross chunk boundaries)
        lens40 = [40, 0, 5, 17, 31, 10, 9, 33, 2, 20, 11]
        want = O.ragged(O.modified_beam_search, m, enc[:B, :40], lens40, 4)
        t, s, sc = h.modified_beam_search(raw[:B, :40], 4, lens=lens40)
        compare_streams(t, s, want, "ragged mbs pipelined", allow_frac=0.15)
    with pytest.raises(_native.K2bError):
        h.greedy_offline(raw[:B, :T], _native.GREEDY_BATCH_COMPAT, lens=lens)      # the reference's coupled loop decodes padding
    with pytest.raises(_native.K2bError):
        h.modified_beam_search(raw[:B - 1, :T], 4, lens=lens)                       # stream count mismatch
    h.close()


def test_cluster16_greedy_online_vocab_2000(built_lib):
    """cfg3 shape (V = 2000): 16-CTA non-portable clusters, greedy online chunks, against the oracle."""
    dims = synth.CONFIGS["cfg3"].dims
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg3"].blank_bias)
    h = make(dims, w, "bf16x3")
    B, Tc = 37, 8
    raw = synth.make_frames(B, 3 * Tc, dims.encoder_dim, 41)
    enc = O.encoder_proj(m, raw)
    hyp = np.zeros((B, 2), np.int64)
    ohyp, otoks = [[0, 0]] * B, [[0, 0]] * B
    h.greedy_online_chunk(np.ascontiguousarray(raw[:, :Tc]), hyp.copy(), enc_is_raw=True)     # builds the 8 GB table
    n0 = h.launch_count()
    for c in range(3):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
        res = O.greedy_search_online_chunk(m, enc[:, Tc * c:Tc * c + Tc], ohyp, otoks)
        ex = compare_streams(t, s, res, f"cluster16 online chunk {c}", allow_frac=0.1)
        if ex:
            break
        ohyp, otoks = [r.hyp for r in res], [r.tokens for r in res]
        assert hyp.tolist() == ohyp
    per_chunk = (h.launch_count() - n0) / 3
    assert per_chunk <= 4, f"{per_chunk} launches per chunk: the 16-CTA cluster path was not taken"
    h.close()


def test_per_frame_tcgen05_joiner_large_vocab(built_lib):
    """cfg4 shape (V = 5537, does not fit a cluster): per-frame path with the tcgen05 joiner (256-column vocab tiles, reducing
    epilogue) for beam search and for greedy, against the oracle."""
    dims = synth.CONFIGS["cfg4"].dims
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg4"].blank_bias)
    h = make(dims, w, "bf16x3")
    raw = synth.make_frames(9, 20, dims.encoder_dim, 43)
    enc = O.encoder_proj(m, raw)
    t, s, sc = h.modified_beam_search(raw, 4, enc_is_raw=True)
    want = O.modified_beam_search(m, enc, 4)
    ex = compare_streams(t, s, want, "per-frame tc joiner mbs V=5537", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=True), "per-frame tc joiner greedy V=5537", allow_frac=0.15)
    h.close()


SHAPES = [
    synth.ModelDims(vocab_size=97, joiner_dim=64, decoder_dim=48, encoder_dim=64),      # one CTA per cluster, 1 k-block
    synth.ModelDims(vocab_size=130, joiner_dim=128, decoder_dim=32, encoder_dim=192),   # 2 slices, last one nearly empty
    synth.ModelDims(vocab_size=1000, joiner_dim=256, decoder_dim=128, encoder_dim=128), # 8 slices, J = 256 (tc encoder_proj)
]


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: f"V{d.vocab_size}_J{d.joiner_dim}_D{d.decoder_dim}_E{d.encoder_dim}")
def test_tensor_paths_over_odd_shapes(built_lib, dims):
    """Every engine of the tcgen05 precisions on shapes that stress the indexing: cluster sizes 1 / 2 / 8, J != 512, D != J,
    vocabulary not a multiple of 128, stream counts that do not fill a cluster, all supported beams."""
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16x3")
    raw = synth.make_frames(11, 17, dims.encoder_dim, 1000 + dims.vocab_size)
    enc = O.encoder_proj(m, raw)
    np.testing.assert_allclose(h.encoder_proj(raw), enc, rtol=0, atol=5e-5)
    for beam in (2, 4, 8, 3):
        if beam > dims.vocab_size:
            continue
        want = O.modified_beam_search(m, enc, beam)
        t, s, sc = h.modified_beam_search(raw, beam, enc_is_raw=True)
        ex = compare_streams(t, s, want, f"{dims.vocab_size}: beam {beam}", allow_frac=0.15)
        for b, r in enumerate(want):
            if b not in ex:
                assert abs(float(sc[b]) - r.score) < SCORE_TOL
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=False), "greedy per_stream", allow_frac=0.15)
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=True), "greedy compat", allow_frac=0.15)
    hyp = np.zeros((11, 2), np.int64)
    t, s, hyp = h.greedy_online_chunk(raw, hyp, enc_is_raw=True)
    res = O.greedy_search_online_chunk(m, enc, [[0, 0]] * 11, [[0, 0]] * 11)
    ex = compare_streams(t, s, res, "online", allow_frac=0.15)
    assert [hyp[b].tolist() for b in range(11) if b not in ex] == [r.hyp for b, r in enumerate(res) if b not in ex]
    h.close()


def test_cluster_beam_merge_with_crafted_collisions(built_lib):
    """blank and unk dominate (both keep ys unchanged) so that merged hypotheses and log-adds occur at every frame."""
    dims = SHAPES[1]
    m, w = model_and_weights(dims)
    w["out_b"] = w["out_b"].copy()
    w["out_b"][0] += 5.0
    w["out_b"][2] += 5.0
    m = O.Model.from_dict(w)
    h = make(dims, w, "bf16x3")
    enc = synth.make_frames(5, 12, dims.joiner_dim, 9)
    for beam in (2, 4):
        want = O.modified_beam_search(m, enc, beam)
        t, s, sc = h.modified_beam_search(enc, beam, enc_is_raw=False)
        ex = compare_streams(t, s, want, f"collisions beam {beam}", allow_frac=0.15)
        np.testing.assert_allclose([sc[b] for b in range(5) if b not in ex], [r.score for b, r in enumerate(want) if b not in ex],
                                   atol=SCORE_TOL)
    h.close()


def test_time_chunk_pipelined_host_call_is_identical(setup):
    """The host-pointer call cuts the utterances into time chunks (H2D of chunk c+1 overlaps projection + search of chunk c,
    hypothesis state carried between cluster-kernel launches): same arithmetic, so bit-identical to the one-launch path
    (forced here by switching the profiler bracket on)."""
    m, w, raw, enc = setup                       # T = 40 >= 32 -> chunks of 8, 24 and 8 frames
    h = make(MID, w, "bf16x3")
    h.modified_beam_search(raw, 4)                # builds the one-off tables / weight images
    for beam in (4, 2):
        n0 = h.launch_count()
        t1, s1, sc1 = h.modified_beam_search(raw, beam)
        n_pipe = h.launch_count() - n0
        h.profile_enable(True)
        n0 = h.launch_count()
        t2, s2, sc2 = h.modified_beam_search(raw, beam)
        n_one = h.launch_count() - n0
        h.profile_enable(False)
        h.profile_read()
        assert t1 == t2 and s1 == s2 and sc1.tolist() == sc2.tolist()
        assert n_one == 3 and n_pipe == 3 * 2 + 1, (n_pipe, n_one)     # proj + cluster (+ per chunk) + back-trace
    compare_streams(t1, s1, O.modified_beam_search(m, enc, 2), "pipelined vs oracle", allow_frac=0.12)
    h.close()


@pytest.mark.parametrize("beam", [2, 4, 8])
def test_persistent_beam_kernel_large_vocab(built_lib, monkeypatch, beam):
    """V = 5537, beams 2 / 4 / 8: the whole time loop in one launch (persistent tcgen05 joiner + merge warps chained by global
    counters) against the oracle, against the per-frame launches (K2B_NO_MEGA=1: identical results required), with more streams
    than fit one row tile, and with ragged lengths."""
    dims = synth.CONFIGS["cfg4"].dims
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg4"].blank_bias)
    h = make(dims, w, "bf16x3")
    B, T = 70, 24                       # 70 * beam rows: 2 .. 5 row tiles, the last one partly filled
    raw = synth.make_frames(B, T, dims.encoder_dim, 77 + beam)
    enc = O.encoder_proj(m, raw)
    t1, s1, sc1 = h.modified_beam_search(raw, beam, enc_is_raw=True)
    h.set_option("no_mega", 1)
    t0, s0, sc0 = h.modified_beam_search(raw, beam, enc_is_raw=True)
    h.set_option("no_mega", 0)
    assert t1 == t0 and s1 == s0
    np.testing.assert_array_equal(np.asarray(sc1), np.asarray(sc0))
    want = O.modified_beam_search(m, enc[:12], beam)
    ex = compare_streams(t1[:12], s1[:12], want, f"persistent beam kernel K={beam}", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc1[b]) - r.score) < SCORE_TOL
    lens = np.array([(7 * b) % (T + 1) for b in range(B)], np.int64)
    h.set_encoder_out_lens(lens)
    t2, s2, _ = h.modified_beam_search(raw, beam, enc_is_raw=True)
    h.set_option("no_mega", 1)
    h.set_encoder_out_lens(lens)
    t3, s3, _ = h.modified_beam_search(raw, beam, enc_is_raw=True)
    assert t2 == t3 and s2 == s3
    assert all(all(x < lens[b] for x in s2[b]) for b in range(B))
    h.close()


@pytest.mark.parametrize("V,J,D,B,T,beam", [
    (1030, 64, 32, 1, 1, 4),          # just above the cluster kernel's vocabulary limit; one stream, one frame
    (1030, 128, 64, 33, 9, 2),        # two k-blocks; 66 rows: one partly filled row tile
    (2500, 64, 48, 70, 7, 8),         # beam 8: 560 rows = 4.4 row tiles, 16 column tiles
    (10300, 64, 32, 5, 6, 4),         # 65 column tiles: more (max, sum) pairs and candidates than a lane keeps in registers
])
def test_persistent_beam_kernel_shapes(built_lib, V, J, D, B, T, beam):
    """Shapes that stress the indexing of the persistent beam kernel and of the merge step (tile counts around the register
    batches, partly filled row tiles, single stream / single frame), against the oracle."""
    dims = synth.ModelDims(vocab_size=V, joiner_dim=J, decoder_dim=D, encoder_dim=64)
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16x3")
    raw = synth.make_frames(B, T, dims.encoder_dim, 5000 + V)
    enc = O.encoder_proj(m, raw)
    t, s, sc = h.modified_beam_search(raw, beam, enc_is_raw=True)
    nb = min(B, 10)
    want = O.modified_beam_search(m, enc[:nb], beam)
    ex = compare_streams(t[:nb], s[:nb], want, f"persistent beam kernel V={V} K={beam}", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
    h.close()


def test_tagged_records_equal_counters_and_survive_engine_changes(built_lib):
    """Persistent large-vocabulary kernel: the joiner CTAs hand their records to the merge warps by epoch tags (default) or through
    release / acquire counters (k2b_set_option("tagged_records", 0)): identical results for beam 4 and for greedy. The partials
    buffer is shared with every other engine and layout: calls are interleaved (beam 4 / greedy / the per-frame launches / the
    fp32 path / beam 8) and every tagged call must still decode what the counter form decodes - a stale word that looked like a
    tag would show here."""
    dims = synth.ModelDims(2100, 128, 64, 64)
    m, w = model_and_weights(dims, blank_bias=0.6)
    h = make(dims, w, "bf16x3")
    B, T = 37, 20
    raw = synth.make_frames(B, T, dims.encoder_dim, 4242)
    ref = {}
    h.set_option("tagged_records", 0)
    ref["b4"] = h.modified_beam_search(raw, 4, enc_is_raw=True)
    ref["b8"] = h.modified_beam_search(raw, 8, enc_is_raw=True)
    ref["g"] = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    h.set_option("tagged_records", 1)

    def same(a, b):
        return a[0] == b[0] and a[1] == b[1] and (len(a) < 3 or np.array_equal(np.asarray(a[2]), np.asarray(b[2])))

    for rnd in range(3):
        assert same(h.modified_beam_search(raw, 4, enc_is_raw=True), ref["b4"]), f"beam 4, round {rnd}"
        assert same(h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True), ref["g"]), f"greedy, round {rnd}"
        h.set_option("no_mega", 1)                                  # per-frame launches write untagged records into the same buffer
        assert same(h.modified_beam_search(raw, 4, enc_is_raw=True), ref["b4"])
        h.set_option("no_mega", 0)
        assert same(h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True), ref["g"]), f"greedy after per-frame launches, round {rnd}"
        h.set_precision("fp32")                                     # the CUDA-core path carves the buffer into four arrays
        h.modified_beam_search(raw[:5], 4, enc_is_raw=True)
        h.set_precision("bf16x3")
        assert same(h.modified_beam_search(raw, 8, enc_is_raw=True), ref["b8"]), f"beam 8 after the fp32 path, round {rnd}"
        assert same(h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True), ref["g"])
    h.close()


def test_persistent_greedy_large_vocab(built_lib, monkeypatch):
    """Greedy search as beam 1 on the persistent beam kernel: V = 5537 (no cluster fits) offline PER_STREAM / SINGLE and online
    chunks against the oracle; V = 2000 online chunks forced onto it (K2B_GREEDY_PERSISTENT=1) against the 16-CTA cluster kernel."""
    dims = synth.CONFIGS["cfg4"].dims
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg4"].blank_bias)
    h = make(dims, w, "bf16x3")
    B, T = 21, 24
    raw = synth.make_frames(B, T, dims.encoder_dim, 91)
    enc = O.encoder_proj(m, raw)
    h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)           # builds the decoder table
    n0 = h.launch_count()
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    assert h.launch_count() - n0 <= 8, "the persistent kernel was not taken"
    compare_streams(t, s, O.greedy_search_batch(m, enc, compat=False), "persistent greedy per_stream V=5537", allow_frac=0.15)
    h.set_option("no_mega", 1)            # the same search as one joiner + one merge launch per frame
    tp, sp = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    h.set_option("no_mega", 0)
    assert tp == t and sp == s
    t1, s1 = h.greedy_offline(raw[:1], _native.GREEDY_SINGLE, enc_is_raw=True)
    ws = O.greedy_search_single(m, enc[0])
    assert (t1[0] == ws.appended and s1[0] == ws.timestamps) or ws.min_gap < 1e-4
    Tc = 8
    hyp = np.zeros((B, 2), np.int64)
    ohyp, otoks = [[0, 0]] * B, [[0, 0]] * B
    for c in range(3):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
        res = O.greedy_search_online_chunk(m, enc[:, Tc * c:Tc * c + Tc], ohyp, otoks)
        ex = compare_streams(t, s, res, f"persistent greedy online chunk {c}", allow_frac=0.15)
        if ex:
            break
        ohyp, otoks = [r.hyp for r in res], [r.tokens for r in res]
        assert hyp.tolist() == ohyp
    h.close()

    dims = synth.CONFIGS["cfg3"].dims
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg3"].blank_bias)
    h = make(dims, w, "bf16x3")
    B = 150
    raw = synth.make_frames(B, 2 * Tc, dims.encoder_dim, 92)
    outs = []
    for force in ("0", "1"):
        h.set_option("greedy_persistent", int(force))
        hyp = np.zeros((B, 2), np.int64)
        got = []
        for c in range(2):
            t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
            got.append((t, s, hyp.tolist()))
        outs.append(got)
    h.set_option("greedy_persistent", -1)
    same = sum(1 for b in range(B) if all(outs[0][c][0][b] == outs[1][c][0][b] and outs[0][c][1][b] == outs[1][c][1][b] for c in range(2)))
    assert same >= B - 2, f"persistent and cluster greedy disagree on {B - same} of {B} streams"      # near ties may differ
    h.close()


def test_beam_one_large_vocab_reports_scores(built_lib):
    """modified_beam_search with beam 1 over a large vocabulary wants the hypothesis score, so it must not take the greedy
    instantiations (which skip the log-softmax): tokens and scores against the oracle."""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=64, decoder_dim=48, encoder_dim=64)
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16x3")
    raw = synth.make_frames(9, 12, dims.encoder_dim, 77)
    enc = O.encoder_proj(m, raw)
    t, s, sc = h.modified_beam_search(raw, 1, enc_is_raw=True)
    want = O.modified_beam_search(m, enc, 1)
    ex = compare_streams(t, s, want, "beam 1 V=2500", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
    tg, sg = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    assert [tg[b] for b in range(9) if b not in ex] == [t[b] for b in range(9) if b not in ex]
    h.close()


@pytest.mark.parametrize("beam", [4, 3])
def test_time_chunked_host_call_large_vocab(built_lib, monkeypatch, beam):
    """Large vocabulary: the host-pointer call steps the persistent beam kernel (beam 4) / the per-frame fused launches (beam 3) in
    time chunks so that the input copy hides behind the search; the hypothesis state is carried in the workspaces. Same arithmetic,
    so bit-identical to the single-chunk call (K2B_PIPE_CHUNKS=1), with ragged lengths too."""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=64, decoder_dim=48, encoder_dim=64)
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16x3")
    B, T = 37, 100                                  # T >= 64: three chunks of 34 frames
    raw = synth.make_frames(B, T, dims.encoder_dim, 123)
    h.modified_beam_search(raw, beam, enc_is_raw=True)
    n0 = h.launch_count()
    t1, s1, sc1 = h.modified_beam_search(raw, beam, enc_is_raw=True)
    n_chunked = h.launch_count() - n0
    h.set_option("pipe_chunks", 1)
    n0 = h.launch_count()
    t2, s2, sc2 = h.modified_beam_search(raw, beam, enc_is_raw=True)
    n_one = h.launch_count() - n0
    h.set_option("pipe_chunks", 0)
    assert t1 == t2 and s1 == s2 and sc1.tolist() == sc2.tolist()
    assert n_chunked > n_one, (n_chunked, n_one)
    lens = [T - (11 * b) % T for b in range(B)]
    t3, s3, _ = h.modified_beam_search(raw, beam, enc_is_raw=True, lens=lens)
    h.set_option("pipe_chunks", 1)
    t4, s4, _ = h.modified_beam_search(raw, beam, enc_is_raw=True, lens=lens)
    assert t3 == t4 and s3 == s4
    enc = O.encoder_proj(m, raw[:6])
    ex = compare_streams(t1[:6], s1[:6], O.modified_beam_search(m, enc, beam), "chunked vs oracle", allow_frac=0.15)
    h.close()


def test_ragged_lengths_large_vocab(built_lib):
    """encoder_out_lens on the persistent kernels (V = 2500: no cluster holds it): beam search and per-stream greedy against the
    oracle's ragged wrapper, lengths 0 and T included."""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=64, decoder_dim=48, encoder_dim=64)
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16x3")
    B, T = 11, 24
    raw = synth.make_frames(B, T, dims.encoder_dim, 55)
    enc = O.encoder_proj(m, raw)
    lens = [24, 0, 5, 17, 1, 24, 9, 13, 2, 20, 7]
    want = O.ragged(O.modified_beam_search, m, enc, lens, 4)
    t, s, sc = h.modified_beam_search(raw, 4, enc_is_raw=True, lens=lens)
    ex = compare_streams(t, s, want, "ragged mbs persistent", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < SCORE_TOL
        assert all(x < lens[b] for x in s[b])
    wantg = O.ragged(O.greedy_search_batch, m, enc, lens, compat=False)
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True, lens=lens)
    compare_streams(t, s, wantg, "ragged greedy persistent", allow_frac=0.15)
    assert all(all(x < lens[b] for x in s[b]) for b in range(B)) and s[1] == [] and t[1] == []
    t2, s2 = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)          # the lengths were consumed by one call
    compare_streams(t2, s2, O.greedy_search_batch(m, enc, compat=False), "after ragged greedy persistent", allow_frac=0.15)
    h.close()


def test_persistent_kernel_bf16_matches_bf16_oracle(built_lib):
    """Single-pass bf16 on the persistent kernels (V = 2500): against the oracle restated with bf16-rounded joiner / encoder_proj
    operands (stated tolerance 5e-3 on the score), beam search and greedy."""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=256, decoder_dim=48, encoder_dim=64)   # J % 256 == 0: tcgen05 encoder_proj too
    m, w = model_and_weights(dims, blank_bias=0.5)
    h = make(dims, w, "bf16")
    raw = synth.make_frames(9, 16, dims.encoder_dim, 66)
    mb = O.Model.from_dict(w, prec_joiner="bf16", prec_enc="bf16")
    encb = O.encoder_proj(mb, raw)
    want = O.modified_beam_search(mb, encb, 4)
    t, s, sc = h.modified_beam_search(raw, 4, enc_is_raw=True)
    ex = compare_streams(t, s, want, "persistent bf16 vs bf16 oracle", allow_frac=0.15)
    for b, r in enumerate(want):
        if b not in ex:
            assert abs(float(sc[b]) - r.score) < 5e-3
    tg, sg = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    compare_streams(tg, sg, O.greedy_search_batch(mb, encb, compat=False), "persistent greedy bf16", allow_frac=0.15)
    h.close()


@pytest.mark.parametrize("a,b", [(10, 50), (10, 100), (10, 300), (1300, 2499)])
def test_persistent_kernel_exact_ties_go_to_the_larger_index(built_lib, a, b):
    """Q1 (ties -> larger index, ref OfflineRecognizer.cs:145-159) on the persistent kernels: two vocabulary rows with identical
    weights and a dominant bias tie exactly in every frame - inside one thread's columns, across the two column halves of a tile,
    across tiles. Greedy must emit the larger id every frame. (Beam search keeps both ids; which hypothesis ends up best depends on
    1e-6 differences of the log-softmax normalisers of the two contexts, so only its score and alphabet are checked.)"""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=64, decoder_dim=48, encoder_dim=64)
    w = synth.make_weights(dims, blank_bias=0.0)
    w["out_w"][b] = w["out_w"][a]
    w["out_b"][a] = w["out_b"][b] = 60.0
    h = make(dims, w, "bf16x3")
    B, T = 5, 9
    raw = synth.make_frames(B, T, dims.encoder_dim, 99)
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    assert all(t[i] == [b] * T and s[i] == list(range(T)) for i in range(B)), t
    tb, sb, scb = h.modified_beam_search(raw, 4, enc_is_raw=True)
    assert all(len(tb[i]) == T and set(tb[i]) <= {a, b} for i in range(B)), tb
    m = O.Model.from_dict(w)
    want = O.modified_beam_search(m, O.encoder_proj(m, raw[:2]), 4)
    for i, r in enumerate(want):
        assert abs(float(scb[i]) - r.score) < SCORE_TOL
    h.close()


def test_persistent_online_masks_id_1(built_lib):
    """The literal `y != 1` of the online loop (ref OnlineRecognizer.cs:181) on the persistent greedy kernel: id 1 dominant in every
    frame is emitted by the offline search and swallowed by the online one, which leaves Hyp untouched."""
    dims = synth.ModelDims(vocab_size=2500, joiner_dim=64, decoder_dim=48, encoder_dim=64)
    w = synth.make_weights(dims, blank_bias=0.0)
    w["out_b"][1] = 60.0
    h = make(dims, w, "bf16x3")
    B, T = 4, 8
    raw = synth.make_frames(B, T, dims.encoder_dim, 98)
    t, s = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    assert all(t[i] == [1] * T for i in range(B))
    hyp = np.array([[7, 9]] * B, np.int64)
    t, s, hyp2 = h.greedy_online_chunk(raw, hyp.copy(), enc_is_raw=True)
    assert all(t[i] == [] and s[i] == [] for i in range(B)) and hyp2.tolist() == hyp.tolist()
    h.close()
