"""GPU (-m gpu): parity at BASELINE.json sizes with the north star's own number (frames bit-exact >= 99.9 %, near ties judged at
the first divergent frame and listed), streaming modified_beam_search, BATCH_COMPAT on the fast engines, the online / CTC sides of
the seam mirror, and the advisor's regressions. Everything goes through the C ABI (ctypes)."""
import json

import numpy as np
import pytest
import torch

from k2transducerasr_b200 import _native, synth
from k2transducerasr_b200 import proj as P
from k2transducerasr_b200 import recognizer as R
from oracle import k2_oracle as O
from tests.helpers import MID, SCORE_TOL, compare_streams, model_and_weights, parity_report

pytestmark = pytest.mark.gpu


def make(dims, w, prec="bf16x3", **kw):
    h = _native.Handle(vocab_size=dims.vocab_size, joiner_dim=dims.joiner_dim, decoder_dim=dims.decoder_dim,
                       encoder_dim=dims.encoder_dim, precision=_native.PREC_NAMES[prec], **kw)
    if w is not None:
        h.load_weights(w)
    return h


def cfg_setup(name, streams=None, frames=None):
    cfg = synth.CONFIGS[name]
    m, w = model_and_weights(cfg.dims, blank_bias=cfg.blank_bias)
    B, T = streams or cfg.streams, frames or cfg.frames * cfg.chunks
    raw = synth.make_frames(B, T, cfg.dims.encoder_dim, cfg.seed)
    return cfg, m, w, raw


# ---- the north star's parity number at the configs' own sizes ---------------------------------------------------------------
def test_parity_full_cfg2_all_streams_all_frames(built_lib):
    """cfg2 exactly as benched: 256 streams x 250 frames, beam 4, V = 500, split-bf16 x3 cluster kernel, raw frames through the
    tcgen05 encoder_proj. Every stream is compared with the oracle - beam history frame by frame, then the final output."""
    cfg, m, w, raw = cfg_setup("cfg2")
    h = make(cfg.dims, w)
    t, s, sc = h.modified_beam_search(raw, cfg.beam)
    bp = h.debug_backpointers(cfg.streams, cfg.frames, cfg.beam)
    want = O.modified_beam_search(m, O.encoder_proj(m, raw), cfg.beam)
    rep = parity_report(t, s, want, scores=sc, bp=bp, T=cfg.frames)
    print("PARITY cfg2", json.dumps(rep.as_dict()))
    rep.assert_ok("cfg2 full", min_frames_pct=99.9)
    assert rep.streams_identical >= int(0.97 * cfg.streams)
    h.close()


def test_parity_full_cfg4_all_streams(built_lib):
    """cfg4: V = 5537, beam 4, 250 frames, persistent beam kernel (one launch for the whole loop); all 256 streams are compared
    with the oracle (about 12 s of CPU on the GPU box)."""
    cfg, m, w, raw = cfg_setup("cfg4")
    h = make(cfg.dims, w)
    t, s, sc = h.modified_beam_search(raw, cfg.beam, enc_is_raw=True)
    bp = h.debug_backpointers(cfg.streams, cfg.frames, cfg.beam)
    n = cfg.streams
    want = O.modified_beam_search(m, O.encoder_proj(m, raw[:n]), cfg.beam)
    rep = parity_report(t[:n], s[:n], want, scores=sc[:n], bp=bp[:n], T=cfg.frames)
    print("PARITY cfg4", json.dumps(rep.as_dict()))
    rep.assert_ok("cfg4 all streams", min_frames_pct=99.9)
    h.close()


def test_parity_full_cfg3_online_512_streams_32_chunks(built_lib):
    """cfg3: 512 concurrent online streams, V = 2000, 32 chunks of 8 frames, Hyp carried across the chunks on our side and the
    oracle's; the whole 256-frame history of every stream is compared."""
    cfg, m, w, raw = cfg_setup("cfg3")
    h = make(cfg.dims, w)
    B, Tc, C = cfg.streams, cfg.frames, cfg.chunks
    enc = O.encoder_proj(m, raw)
    hyp = np.zeros((B, 2), np.int64)
    got_t, got_s = [[] for _ in range(B)], [[] for _ in range(B)]
    for c in range(C):
        t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
        for b in range(B):
            got_t[b] += t[b]
            got_s[b] += [x + Tc * c for x in s[b]]
    # online greedy over the chunks == one long chunk (Q6 is a no-op online): one oracle call, absolute timestamps
    want = O.greedy_search_online_chunk(m, enc, [[0, 0]] * B, [[0, 0]] * B)
    rep = parity_report(got_t, got_s, want, T=Tc * C)
    print("PARITY cfg3", json.dumps(rep.as_dict()))
    rep.assert_ok("cfg3 full", min_frames_pct=99.9)
    same_hyp = sum(1 for b in range(B) if hyp[b].tolist() == want[b].hyp)
    assert same_hyp >= rep.streams_identical
    h.close()


def test_online_chunks_projected_ahead_on_a_side_stream(built_lib):
    """cfg3 shape through the device-pointer entry with "inputs_complete": the encoder_proj of chunk c+1 runs on a side stream
    under the persistent search of chunk c (which leaves 20 SMs free), projected frames in two buffers taken in turns. Chunks
    enqueued back to back without a host synchronisation must give what the stream-ordered form gives, bit for bit - also with the
    host-pointer entry (whose staged input is NOT complete at call time) in between."""
    cfg, m, w, raw = cfg_setup("cfg3")
    h = make(cfg.dims, w)
    B, Tc, C = cfg.streams, cfg.frames, 10
    chunks = [torch.from_numpy(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc])).cuda() for c in range(C)]
    outs = {}
    for complete in (0, 1):
        h.set_option("inputs_complete", complete)
        hyp = torch.zeros((B, 2), dtype=torch.int64, device="cuda")
        tok = torch.zeros((C, B, Tc), dtype=torch.int64, device="cuda"); ts = torch.zeros((C, B, Tc), dtype=torch.int32, device="cuda")
        n = torch.zeros((C, B), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for c in range(C):
            if c == 5:        # a host-pointer chunk in the middle (same frames, host Hyp round trip)
                hh = hyp.cpu().numpy()
                t5, s5, hh = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hh, enc_is_raw=True)
                hyp.copy_(torch.from_numpy(hh))
                for b in range(B):
                    n[c, b] = len(t5[b])
                    if t5[b]:
                        tok[c, b, :len(t5[b])] = torch.tensor(t5[b]); ts[c, b, :len(t5[b])] = torch.tensor(s5[b], dtype=torch.int32)
                continue
            h.call("k2b_greedy_online_chunk_dev", chunks[c], 1, B, Tc, hyp, tok[c], ts[c], n[c], Tc)
        h.sync()
        outs[complete] = (tok, ts, n, hyp)
    h.set_option("inputs_complete", 0)
    a, b = outs[0], outs[1]
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and int(a[2].sum()) > 0
    mask = torch.arange(Tc, device="cuda")[None, None, :] < a[2][:, :, None]
    assert torch.equal(a[0][mask], b[0][mask]) and torch.equal(a[1][mask], b[1][mask])
    h.close()


def test_parity_full_cfg1_single_stream(built_lib):
    """cfg1: one utterance of 250 frames, offline single-stream greedy (the reference's own CPU-runnable case)."""
    cfg, m, w, raw = cfg_setup("cfg1")
    h = make(cfg.dims, w)
    t, s = h.greedy_offline(raw, _native.GREEDY_SINGLE)
    want = [O.greedy_search_single(m, O.encoder_proj(m, raw)[0])]
    rep = parity_report(t, s, want, T=cfg.frames)
    print("PARITY cfg1", json.dumps(rep.as_dict()))
    rep.assert_ok("cfg1", min_frames_pct=99.0)          # one stream: a single near-tie frame may cost the rest of it
    h.close()


def test_parity_full_cfg5_ctc(built_lib):
    """cfg5 at one GPU's share (128 streams x 250 frames, V = 2000): bit-exact, no excuses (integer / index work)."""
    cfg = synth.CONFIGS["cfg5"]
    lp = synth.make_ctc_logp(128, cfg.frames, cfg.dims.vocab_size, cfg.seed, blank_bias=cfg.blank_bias)
    h = make(cfg.dims, None, "fp32")
    t, s, tb, _ = h.ctc_greedy(lp, trailing_blank=np.zeros(128, np.int32))
    want = O.ctc_greedy_search(lp)
    assert t == [r.appended for r in want] and s == [r.timestamps for r in want]
    assert tb.tolist() == [r.num_trailing_blank for r in want]
    h.close()


# ---- streaming modified_beam_search ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cluster_v500", "persistent_v2000", "fp32_v500", "beam3_v500"])
def test_streaming_beam_search_chunked_equals_whole(built_lib, case):
    """Hypotheses carried between chunks in device slots: decoding chunk by chunk (uneven chunks, streams joining late, slots in
    arbitrary order) == the oracle's whole-utterance beam search with the online seed and mask, and - bit for bit, scores
    included - == our own single-chunk call."""
    if case == "persistent_v2000":
        dims, bias, prec, beam = synth.CONFIGS["cfg3"].dims, synth.CONFIGS["cfg3"].blank_bias, "bf16x3", 4
    else:
        dims, bias, prec, beam = MID, 0.99, ("fp32" if case == "fp32_v500" else "bf16x3"), (3 if case == "beam3_v500" else 4)
    m, w = model_and_weights(dims, blank_bias=bias)
    B, T = 13, 40
    raw = synth.make_frames(B, T, dims.encoder_dim, 808)
    enc = O.encoder_proj(m, raw)
    h = make(dims, w, prec)
    h.beam_pool_create(32, beam, 64)
    slots = [3 * b % 32 for b in range(B)]
    for sl in slots:
        h.beam_pool_reset(sl, [0, 0])
    # all streams, uneven chunks
    off = 0
    for n in (8, 8, 1, 16, 7):
        t, s, sc, hyp = h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:, off:off + n]), slots, cap=64, enc_is_raw=True)
        off += n
    want = O.modified_beam_search(m, enc, beam, init=O.mbs_seed(m, B, [[0, 0]] * B), extra_mask=1)
    ex = compare_streams(t, s, want, f"streaming {case}", allow_frac=0.2, scores=sc, T=T)
    for b, r in enumerate(want):
        if b not in ex:
            assert hyp[b].tolist() == r.hyp
    # single chunk in fresh slots: identical to the chunked run, bit for bit
    slots2 = [31 - b for b in range(B)]
    for sl in slots2:
        h.beam_pool_reset(sl, [0, 0])
    t1, s1, sc1, hyp1 = h.modified_beam_search_online_chunk(raw, slots2, cap=64, enc_is_raw=True)
    assert t1 == t and s1 == s and hyp1.tolist() == hyp.tolist()
    np.testing.assert_array_equal(sc1, sc)
    # a stream that joins late and a reset in the middle: slots are independent
    for sl in (0, 1):
        h.beam_pool_reset(sl, [0, 0])
    h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:1, :10]), [0], cap=64, enc_is_raw=True)
    ta, sa, sca, _ = h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:2, 10:20]), [0, 1], cap=64, enc_is_raw=True)
    h.beam_pool_reset(2, [0, 0])
    tb, sb, scb, _ = h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:1, :20]), [2], cap=64, enc_is_raw=True)
    assert ta[0] == tb[0] and sa[0] == sb[0] and sca[0] == scb[0]
    h.beam_pool_reset(3, [0, 0])
    tc, sc_, scc, _ = h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[1:2, 10:20]), [3], cap=64, enc_is_raw=True)
    assert ta[1] == tc[0] and [x for x in sa[1]] == sc_[0] and sca[1] == scc[0]
    # errors: the same slot twice, more frames than the pool holds
    with pytest.raises(_native.K2bError):
        h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:2, :4]), [5, 5], cap=64, enc_is_raw=True)
    with pytest.raises(_native.K2bError):
        h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:1]), [slots2[5]], cap=64, enc_is_raw=True)   # 40 + 40 > 64
    h.close()


def test_online_recognizer_modified_beam_search_and_endpoint(built_lib):
    """OnlineRecognizer(decodingMethod="modified_beam_search", maxActivePaths=4) through the mirror of the reference API: the result
    after the last chunk equals the oracle's whole-utterance search; NumTrailingBlank / IsEndpoint / Reset behave."""
    dims = MID
    m, w = model_and_weights(dims, blank_bias=0.99)
    proj = P.OnlineProjOfB200(dims, w, chunk_frames=8, precision="bf16x3")
    rec = R.OnlineRecognizer(proj, decodingMethod="modified_beam_search", maxActivePaths=4, enableEndpoint=1, maxStreams=8, maxFrames=256)
    B, T = 3, 32
    raw = synth.make_frames(B, T, dims.encoder_dim, 4711)
    streams = [rec.CreateOnlineStream() for _ in range(B)]
    for b, st in enumerate(streams):
        st.AcceptFrames(raw[b])
    for _ in range(T // 8):
        rec.GetResults(list(streams))
    want = O.modified_beam_search(m, O.encoder_proj(m, raw), 4, init=O.mbs_seed(m, B, [[0, 0]] * B), extra_mask=1)
    compare_streams([st.Tokens[2:] for st in streams], [st.Timestamps for st in streams], want, "online recognizer mbs", allow_frac=0.34, T=T)
    for st, r in zip(streams, want):
        assert st.NumProcessedFrames == T
        if st.Tokens[2:] == r.appended:
            assert st.NumTrailingBlank == (T - 1 - r.timestamps[-1] if r.timestamps else T)
            assert st.Hyp.tolist() == r.hyp
    # silence: frames that decode to blank push NumTrailingBlank over rule 2 / rule 1
    st = streams[0]
    before = (list(st.Tokens), list(st.Timestamps))
    st.NumTrailingBlank = rec.RULE1_FRAMES
    assert rec.IsEndpoint(st)
    rec.Reset(st)
    assert st.Tokens == [0, 0] and st.NumTrailingBlank == 0 and not rec.IsEndpoint(st)
    st.AcceptFrames(raw[0])
    for _ in range(T // 8):
        rec.GetResults([st])
    assert (list(st.Tokens), list(st.Timestamps)) == before          # a reset stream decodes like a fresh one
    for s_ in streams:
        rec.ReleaseOnlineStream(s_)
    rec.Dispose()


# ---- BATCH_COMPAT on the fast engines (Q5 / Q6) --------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cluster_v500", "persistent_v5537"])
def test_batch_compat_on_fast_engines(built_lib, case):
    """The reference's batch loop (decoder refresh for ALL streams when ANY emits, Q6) as two per-stream passes on the persistent
    kernels: equals the oracle's literal restatement and the three-launches-per-frame path; a batch in which one stream emits
    first and the others much later (the case Q6 is about); a batch that never emits."""
    if case == "cluster_v500":
        dims, bias = MID, 0.99
    else:
        dims, bias = synth.CONFIGS["cfg4"].dims, synth.CONFIGS["cfg4"].blank_bias
    m, w = model_and_weights(dims, blank_bias=bias)
    B, T = 19, 40
    raw = synth.make_frames(B, T, dims.encoder_dim, 606)
    enc = O.encoder_proj(m, raw)
    h = make(dims, w)
    h.greedy_offline(raw[:2, :4], _native.GREEDY_PER_STREAM, enc_is_raw=True)      # builds the weight images and the decoder table
    n0 = h.launch_count()
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)
    assert h.launch_count() - n0 <= 16, "BATCH_COMPAT did not run on the fast engines"
    want = O.greedy_search_batch(m, enc, compat=True)
    compare_streams(t, s, want, f"compat fast {case}", allow_frac=0.12, coupled=True, T=T)
    tp, sp = h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)
    wantp = O.greedy_search_batch(m, enc, compat=False)
    compare_streams(tp, sp, wantp, f"per-stream {case}", allow_frac=0.12, T=T)
    # Q6 must be visible in this batch: some stream decodes differently in the two modes
    assert any(a.appended != b.appended for a, b in zip(want, wantp)) or all(r.appended == [] for r in want)
    h.set_precision("fp32")                            # the per-frame launches implement the loop literally
    t32, s32 = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)
    same = sum(1 for b in range(B) if t32[b] == t[b] and s32[b] == s[b])
    assert same >= B - 2, f"fast and per-frame BATCH_COMPAT disagree on {B - same} streams"
    h.close()
    # a joiner that only ever says blank: nobody emits, nothing to redo
    w2 = dict(w)
    w2["out_b"] = w["out_b"].copy()
    w2["out_b"][0] += 50.0
    h = make(dims, w2)
    t, s = h.greedy_offline(raw[:5], _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)
    assert t == [[]] * 5
    h.close()


def test_batch_compat_cfg2_size(built_lib):
    """BATCH_COMPAT at a config-size batch (256 streams x 250 frames, V = 500) on the cluster kernel against the oracle."""
    cfg, m, w, raw = cfg_setup("cfg2")
    h = make(cfg.dims, w)
    t, s = h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT)
    want = O.greedy_search_batch(m, O.encoder_proj(m, raw), compat=True)
    rep = parity_report(t, s, want, coupled=True, T=cfg.frames)
    print("PARITY cfg2-size BATCH_COMPAT", json.dumps(rep.as_dict()))
    rep.assert_ok("compat cfg2 size", min_frames_pct=99.5)
    h.close()


# ---- advisor regressions --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["bf16x3", "fp32"])
def test_pending_lens_do_not_leak_into_online_chunks(built_lib, prec):
    """k2b_set_encoder_out_lens followed by an online chunk must equal the call without it (and must not read out of
    bounds when the online batch is larger than the lens array)."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w, prec)
    raw = synth.make_frames(9, 8, MID.encoder_dim, 55)
    hyp0 = np.zeros((9, 2), np.int64)
    a = h.greedy_online_chunk(raw, hyp0, enc_is_raw=True)
    h.set_encoder_out_lens([0, 0, 0])
    b = h.greedy_online_chunk(raw, hyp0, enc_is_raw=True)
    assert a[0] == b[0] and a[1] == b[1] and a[2].tolist() == b[2].tolist()
    t1, _, _ = h.modified_beam_search(raw, 4)                  # and the setting is gone: all 8 frames are decoded
    assert sum(len(x) for x in t1) > 0
    h.close()


def test_hyp_ids_are_validated(built_lib):
    """caller-supplied Hyp ids outside the vocabulary are rejected (host pointers) or decoded as blank and reported
    by k2b_sync (device pointers) - never used as table addresses."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    raw = synth.make_frames(4, 8, MID.encoder_dim, 56)
    for bad in ([[0, MID.vocab_size]], [[-2, 0]], [[0, -1]]):
        hyp = np.zeros((4, 2), np.int64)
        hyp[2] = bad[0]
        with pytest.raises(_native.K2bError) as e:
            h.greedy_online_chunk(raw, hyp, enc_is_raw=True)
        assert e.value.status == _native.K2B_ERR_INVALID
    dev = torch.from_numpy(raw).cuda()
    hyp = torch.zeros((4, 2), dtype=torch.int64, device="cuda")
    hyp[1, 1] = 10 ** 9
    tok = torch.zeros((4, 8), dtype=torch.int64, device="cuda"); ts = torch.zeros((4, 8), dtype=torch.int32, device="cuda")
    n = torch.zeros(4, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    h.call("k2b_greedy_online_chunk_dev", dev, 1, 4, 8, hyp, tok, ts, n, 8)
    with pytest.raises(_native.K2bError) as e:
        h.sync()
    assert e.value.status == _native.K2B_ERR_INVALID
    h.sync()                                                   # reported once
    good = h.greedy_online_chunk(raw, np.zeros((4, 2), np.int64), enc_is_raw=True)
    assert good[0][1] == tok[1, :int(n[1])].tolist()           # the bad Hyp was decoded as {blank, blank}
    h.close()


def test_decoder_table_that_does_not_fit_falls_back(built_lib):
    """when the memoised decoder table cannot be allocated the tcgen05 precisions take the per-frame path instead of
    failing (and nothing leaks per retry). Simulated by occupying the device memory first."""
    dims = synth.CONFIGS["cfg3"].dims                            # table 8.2 GB
    m, w = model_and_weights(dims, blank_bias=synth.CONFIGS["cfg3"].blank_bias)
    free, total = torch.cuda.mem_get_info()
    hog = torch.empty(int(free - (9 << 30)), dtype=torch.uint8, device="cuda")       # leaves 9 GB: 60 % of it is less than the table
    try:
        h = make(dims, w)
        raw = synth.make_frames(6, 8, dims.encoder_dim, 57)
        before = torch.cuda.mem_get_info()[0]
        t, s, sc = h.modified_beam_search(raw, 4, enc_is_raw=True)
        assert h.get_stat("decoder_table_state") == -1
        want = O.modified_beam_search(m, O.encoder_proj(m, raw), 4)
        compare_streams(t, s, want, "no table: per-frame path", allow_frac=0.34, scores=sc, T=8)
        g, gs, _ = h.greedy_online_chunk(raw, np.zeros((6, 2), np.int64), enc_is_raw=True)
        wg = O.greedy_search_online_chunk(m, O.encoder_proj(m, raw), [[0, 0]] * 6, [[0, 0]] * 6)
        compare_streams(g, gs, wg, "no table: greedy online on the per-frame path", allow_frac=0.34, T=8)
        for _ in range(3):
            h.modified_beam_search(raw, 4, enc_is_raw=True)
        assert before - torch.cuda.mem_get_info()[0] < (1 << 30)
        h.close()
    finally:
        del hog
        torch.cuda.empty_cache()


# ---- f1: more than one symbol per frame ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_max_sym_per_frame(built_lib, prec):
    """max_sym_per_frame > 1 (ref OfflineRecognizer.cs:19, :127-134 - fixed to 1 there): a frame is re-evaluated with the refreshed
    decoder until it yields blank or the bound is met. Raw regime (no blank bias) so that frames do emit several symbols."""
    m, w = model_and_weights(MID, blank_bias=0.0)
    raw = synth.make_frames(5, 12, MID.encoder_dim, 909)
    enc = O.encoder_proj(m, raw)
    h = make(MID, w, prec)
    for msf in (2, 3):
        h.set_option("max_sym_per_frame", msf)
        cap = 12 * msf
        tokens = np.zeros((5, cap), np.int64); ts = np.zeros((5, cap), np.int32); n = np.zeros(5, np.int32)
        h.call("k2b_greedy_offline", np.ascontiguousarray(enc), 0, 5, 12, _native.GREEDY_PER_STREAM, tokens, ts, n, cap)
        for b in range(5):
            r = O.greedy_search_single(m, enc[b], max_sym_per_frame=msf)
            got_t, got_s = tokens[b, :n[b]].tolist(), ts[b, :n[b]].tolist()
            assert (got_t == r.appended and got_s == r.timestamps) or r.min_gap < 1e-4, (msf, b)
        assert n.max() > 12, "no frame emitted more than one symbol: the case is not exercised"
    h.set_option("max_sym_per_frame", 1)
    t, s = h.greedy_offline(enc, _native.GREEDY_PER_STREAM, enc_is_raw=False)
    assert all(len(set(x)) == len(x) for x in s)
    h.close()


# ---- the online / CTC sides of the seam mirror ----------------------------------------------------------------------------------
def test_online_recognizer_fused_on_gpu_with_device_states(built_lib):
    """OnlineRecognizer(fused=True) over OnlineProjOfB200 on the GPU: chunked greedy equals the fine-grained loop and the oracle;
    the encoder caches live in the device pool (GetEncoderInitStates / stack_states / unstack_states = k2b_state_pool_*), and a
    stand-in encoder that adds 1 to every cache value per chunk shows up as the chunk count in every stream's slot."""
    dims = MID
    m, w = model_and_weights(dims, blank_bias=0.99)
    layout = P.zipformer2_state_layout((2, 2), (192, 256), (4, 4), (32, 32), (12, 12), (64, 32), (31, 31))

    def fake_encoder(buf):
        buf += 1.0
        torch.cuda.synchronize()

    proj = P.OnlineProjOfB200(dims, w, chunk_frames=8, precision="bf16x3", state_layout=layout, max_streams=8, encoder_hook=fake_encoder)
    projf = P.OnlineProjOfB200(dims, w, chunk_frames=8, precision="fp32")
    rec = R.OnlineRecognizer(proj, fused=True)
    recf = R.OnlineRecognizer(projf, fused=False)
    B, T = 4, 24
    raw = synth.make_frames(B, T, dims.encoder_dim, 31337)
    ss, sf = [rec.CreateOnlineStream() for _ in range(B)], [recf.CreateOnlineStream() for _ in range(B)]
    assert isinstance(ss[0].States, P.DeviceStates) and len({s.States.slot for s in ss}) == B
    for b in range(B):
        ss[b].AcceptFrames(raw[b, : (T if b else 16)])       # stream 0 has one chunk less: it drops out of the last batch
        sf[b].AcceptFrames(raw[b, : (T if b else 16)])
    for _ in range(3):
        rec.GetResults(list(ss))
        recf.GetResults(list(sf))
    enc = O.encoder_proj(m, raw)
    for b in range(B):
        nfr = T if b else 16
        hyps, toks, tss, gap = [[0, 0]], [[0, 0]], [], float("inf")
        for c in range(nfr // 8):
            r = O.greedy_search_online_chunk(m, enc[b:b + 1, 8 * c:8 * c + 8], hyps, toks)[0]
            hyps, toks = [r.hyp], [r.tokens]
            tss += r.timestamps
            gap = min(gap, r.min_gap)
        assert (ss[b].Tokens == sf[b].Tokens and ss[b].Timestamps == sf[b].Timestamps) or gap < 1e-4
        assert (sf[b].Tokens == toks[0] and sf[b].Timestamps == tss) or gap < 1e-4
        assert (ss[b].Tokens == toks[0] and ss[b].Timestamps == tss) or gap < 1e-4
        got = proj.Native.state_pool_get(ss[b].States.slot)
        np.testing.assert_array_equal(got, np.full_like(got, nfr // 8))
    for s_ in ss:
        rec.ReleaseOnlineStream(s_)
    rec.Dispose(); recf.Dispose()


def test_ctc_projs_through_the_recognizers(built_lib):
    """OfflineProjOfB200ctc / OnlineProjOfB200ctc (ref OfflineProjOfZipformer2ctc.cs, OnlineProjOfZipformer2ctc.cs): DecoderProj /
    JoinerProj return null (ref :93-101), model_type forces greedy_search_ctc (ref OfflineRecognizer.cs:46-49), and the three CTC
    loops leave what the oracle's restatement leaves - incl. Q10 (online: prev_id reset per chunk, nothing written back)."""
    dims = synth.ModelDims(vocab_size=200, joiner_dim=64, decoder_dim=64, encoder_dim=0)
    lp = synth.make_ctc_logp(3, 24, 200, 77, blank_bias=4.0)
    lp[2, 7, :] = -9.0; lp[2, 7, 11] = -0.1
    lp[2, 8, :] = -9.0; lp[2, 8, 11] = -0.1                     # token 11 on both sides of the chunk edge at frame 8 (Q10)
    off = P.OfflineProjOfB200ctc(dims)
    assert off.DecoderProj(None, 1) is None and off.JoinerProj(None, None) is None
    rec = R.OfflineRecognizer(off, decodingMethod="greedy_search")
    want = O.ctc_greedy_search(lp)
    s1 = rec.CreateOfflineStream(); s1.AcceptFrames(lp[0])
    rec.GetResult(s1)
    assert s1.Tokens == [-1, 0] + want[0].appended and s1.Timestamps == want[0].timestamps
    assert s1.NumTrailingBlank == want[0].num_trailing_blank
    ss = [rec.CreateOfflineStream() for _ in range(3)]
    for b, s_ in enumerate(ss):
        s_.AcceptFrames(lp[b])
    rec.GetResults(ss)
    for b, s_ in enumerate(ss):
        assert s_.Tokens == [0, 0] + want[b].appended and s_.Timestamps == want[b].timestamps
        assert s_.NumTrailingBlank == want[b].num_trailing_blank
    rec.Dispose()
    on = P.OnlineProjOfB200ctc(dims, chunk_frames=8)
    assert on.DecoderProj(None, 1) is None and on.JoinerProj(None, None) is None
    reco = R.OnlineRecognizer(on, decodingMethod="greedy_search")
    so = [reco.CreateOnlineStream() for _ in range(3)]
    for b, s_ in enumerate(so):
        s_.AcceptFrames(lp[b])
    for _ in range(3):
        reco.GetResults(list(so))
    for b, s_ in enumerate(so):
        toks, tss = [0, 0], []
        for c in range(3):
            r = O.ctc_greedy_search(lp[b:b + 1, 8 * c:8 * c + 8])[0]          # prev resets, frame_offset stays 0 (Q10)
            toks += r.appended
            tss += r.timestamps
        assert s_.Tokens == toks and s_.Timestamps == tss
        assert s_.FrameOffset == 0 and s_.NumTrailingBlank == 0
    assert so[2].Tokens.count(11) >= 2                                        # emitted on both sides of the edge
    reco.Dispose()


# ---- f4: hot words in the merge step --------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cluster_bf16x3", "perframe_fp32", "large_vocab_unfused"])
def test_hotword_biasing_in_the_merge(built_lib, case):
    """k2b_set_context_graph: the boost of a dense hot-word automaton applied inside the beam search's merge step (cluster kernel and
    per-frame merge), against the oracle's biased search - offline (finalised), through the time-chunked host call, and streaming
    (state carried in the beam pool, not finalised); other engines answer UNSUPPORTED; clearing the graph restores the plain search."""
    from k2transducerasr_b200.hotwords import ContextGraph
    if case == "large_vocab_unfused":
        dims, bias, prec = synth.ModelDims(2100, 64, 32, 64), 0.5, "bf16x3"
    else:
        dims, bias, prec = MID, 0.99, ("fp32" if case == "perframe_fp32" else "bf16x3")
    m, w = model_and_weights(dims, blank_bias=bias)
    B, T, K = 12, 40, 4
    raw = synth.make_frames(B, T, dims.encoder_dim, 1234)
    enc = O.encoder_proj(m, raw)
    plain = O.modified_beam_search(m, enc, K)
    seen = [t for r in plain for t in r.appended]
    words = [seen[0:2], seen[2:5], seen[6:7], seen[9:12], [dims.vocab_size - 1, dims.vocab_size - 2]]
    g = ContextGraph.build(words, dims.vocab_size, 2.0)
    want = O.modified_beam_search(m, enc, K, context_graph=g)
    assert any(a.appended != b.appended or abs(a.score - b.score) > 1.0 for a, b in zip(plain, want))
    h = make(dims, w, prec)
    h.set_context_graph(g)
    if case == "large_vocab_unfused":
        with pytest.raises(_native.K2bError) as e:
            h.modified_beam_search(raw, K, enc_is_raw=True)
        assert e.value.status == _native.K2B_ERR_UNSUPPORTED
        h.set_option("unfused_step", 1)
    t, s, sc = h.modified_beam_search(raw, K, enc_is_raw=True)                # T >= 32 raw frames: the time-chunked host call
    bp = h.debug_backpointers(B, T, K)
    compare_streams(t, s, want, f"hot words {case}", allow_frac=0.2, scores=sc, bp=bp, T=T)
    t2, s2, sc2 = h.modified_beam_search(enc, K, enc_is_raw=False)            # projected frames
    compare_streams(t2, s2, want, f"hot words {case} projected", allow_frac=0.2, scores=sc2, T=T)
    if case != "large_vocab_unfused":
        # streaming: the automaton state travels with the hypotheses; results "so far" keep the boost of a match in progress
        h.beam_pool_create(B, K, 64)
        for sl in range(B):
            h.beam_pool_reset(sl, None)                                       # offline seed {-1, blank}: same search as above
        for lo, hi in ((0, 8), (8, 9), (9, 40)):
            ts_, ss_, scs_, _ = h.modified_beam_search_online_chunk(np.ascontiguousarray(raw[:, lo:hi]), list(range(B)), cap=64, enc_is_raw=True)
        want_s = O.modified_beam_search(m, enc, K, context_graph=g, finalize=False, extra_mask=1)
        compare_streams(ts_, ss_, want_s, f"hot words {case} streaming", allow_frac=0.2, scores=scs_, T=T)
    # greedy search is not biased; clearing the graph restores the plain beam search
    gt, gs = h.greedy_offline(enc, _native.GREEDY_PER_STREAM, enc_is_raw=False)
    compare_streams(gt, gs, O.greedy_search_batch(m, enc, compat=False), f"hot words {case}: greedy untouched", allow_frac=0.2, T=T)
    h.set_context_graph(None)
    t3, s3, sc3 = h.modified_beam_search(raw, K, enc_is_raw=True)
    compare_streams(t3, s3, plain, f"hot words {case} cleared", allow_frac=0.2, scores=sc3, T=T)
    h.close()


# ---- host memory, overlap, multi-GPU gather, diagnostics ---------------------------------------------------------------------
def test_pinned_host_buffers_and_async_results(built_lib):
    """k2b_host_alloc'ed frame / result buffers and "async_d2h": two batches in flight on one handle, completed by k2b_sync,
    equal the synchronous calls."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    B, T, K = 24, 40, 4
    raws = [synth.make_frames(B, T, MID.encoder_dim, 100 + i) for i in range(2)]
    ref = [h.modified_beam_search(r, K) for r in raws]
    pin = [_native.Handle.host_alloc((B, T, MID.encoder_dim), np.float32) for _ in range(2)]
    outs = [(_native.Handle.host_alloc((B, T), np.int64), _native.Handle.host_alloc((B, T), np.int32),
             _native.Handle.host_alloc((B,), np.int32), _native.Handle.host_alloc((B,), np.float32)) for _ in range(2)]
    h.set_option("async_d2h", 1)
    for i in range(2):
        pin[i][...] = raws[i]
        tok, ts, n, sc = outs[i]
        h.call("k2b_modified_beam_search", pin[i], 1, B, T, K, tok, ts, n, sc, T)
    h.sync()
    h.set_option("async_d2h", 0)
    for i in range(2):
        tok, ts, n, sc = outs[i]
        got_t, got_s = _native.Handle._unpack(tok, ts, n)
        assert got_t == ref[i][0] and got_s == ref[i][1]
        np.testing.assert_array_equal(sc, ref[i][2])
    for a in pin + [x for o in outs for x in o]:
        _native.Handle.host_free(a)
    reg = np.ascontiguousarray(raws[0])
    assert _native.lib().k2b_host_register(reg.ctypes.data, reg.nbytes) == 0
    t, s, sc = h.modified_beam_search(reg, K)
    assert t == ref[0][0]
    assert _native.lib().k2b_host_unregister(reg.ctypes.data) == 0
    # pageable arrays are staged by the library's own copy threads: any pool size (0 = the calling thread copies) gives the same
    for th in (0, 1, 3, -1):
        h.set_option("copy_threads", th)
        for i in range(2):
            t, s, sc = h.modified_beam_search(raws[i], K)
            assert t == ref[i][0] and s == ref[i][1]
    h.close()


def test_search_polling_for_frames_projected_on_a_side_stream(built_lib):
    """Device-pointer beam search on raw frames, dev_chunks = 3 (the default for a batch that fills the cluster kernel): the frames
    are projected in time chunks of 32 on a side stream while ONE cluster-kernel launch searches them, polling a per-chunk flag in
    front of each chunk's first frame. With "inputs_complete" the side stream is not ordered behind the handle's stream (the next
    call's chunks are projected under the current search, two projected-frame buffers in turns). Calls that follow each other
    without a host synchronisation - different inputs, an odd frame count, ragged lengths, another engine and a host-pointer call
    in between, a batch that grows - must equal the plain one-launch form bit for bit."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    K = 4

    def run(raws, T, lens=None):
        res = []
        for x in raws:
            B = x.shape[0]
            tok = torch.zeros((B, T), dtype=torch.int64, device="cuda"); ts = torch.zeros((B, T), dtype=torch.int32, device="cuda")
            n = torch.zeros(B, dtype=torch.int32, device="cuda"); sc = torch.zeros(B, dtype=torch.float32, device="cuda")
            if lens is not None:
                h.set_encoder_out_lens(lens[:B])
            h.call("k2b_modified_beam_search_dev", x, 1, B, T, K, tok, ts, n, sc, T)
            res.append((tok, ts, n, sc))
        h.sync()
        return res

    for T, ragged in ((64, False), (77, False), (64, True)):
        raws = [torch.from_numpy(synth.make_frames(128, T, MID.encoder_dim, 900 + i)).cuda() for i in range(5)]
        raws.append(torch.from_numpy(synth.make_frames(160, T, MID.encoder_dim, 990)).cuda())      # the buffers grow
        lens = np.random.default_rng(T).integers(T // 2, T + 1, size=160).astype(np.int64) if ragged else None
        torch.cuda.synchronize()
        h.set_option("inputs_complete", 0); h.set_option("dev_chunks", 1)
        want = run(raws, T, lens)
        for complete, L in ((0, 32), (1, 32), (1, 7), (0, 0), (1, 0)):          # L = 0: chosen by the library (whole GEMM waves)
            h.set_option("dev_chunks", 3 if complete else -1); h.set_option("inputs_complete", complete)
            h.set_option("dev_chunk_frames", L)
            n0 = h.launch_count()
            got = run(raws, T, lens)
            # per call: one GEMM + one flag kernel per chunk of L frames, ONE search launch, the back-trace
            assert L == 0 or h.launch_count() - n0 == len(raws) * (2 * ((T + L - 1) // L) + 2), (h.launch_count() - n0, T, L)
            # another engine and a host-pointer call between such calls
            g1 = run(raws[:1], T, lens)
            h.greedy_offline(synth.make_frames(4, 24, MID.encoder_dim, 1), 1)
            h.modified_beam_search(synth.make_frames(24, 40, MID.encoder_dim, 2), K)
            g2 = run(raws[1:3], T, lens)
            for a, b in zip(want + want[:3], got + g1 + g2):
                for x, y in zip(a, b):
                    assert torch.equal(x, y), (T, ragged, complete, L)
        h.set_option("dev_chunk_frames", 0)
    h.set_option("inputs_complete", 0)
    h.close()


def test_nccl_gather_single_rank(built_lib):
    """k2b_nccl_unique_id / k2b_nccl_init / k2b_gather_results_nccl with one rank: the gather is the identity (the N-rank case runs
    in bench.py under torchrun; here the dlopen'ed libnccl, the communicator and the stream ordering are exercised)."""
    import ctypes as C
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    uid = (C.c_char * 128)()
    assert _native.lib().k2b_nccl_unique_id(uid) == 0
    h.call("k2b_nccl_init", C.addressof(uid), 0, 1)
    B, T, K = 8, 16, 4
    raw = torch.from_numpy(synth.make_frames(B, T, MID.encoder_dim, 5)).cuda()
    tok = torch.zeros((B, T), dtype=torch.int64, device="cuda"); ts = torch.zeros((B, T), dtype=torch.int32, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda"); sc = torch.zeros(B, dtype=torch.float32, device="cuda")
    atok, ats, an, asc = torch.full_like(tok, -1), torch.full_like(ts, -1), torch.full_like(n, -1), torch.full_like(sc, -1)
    torch.cuda.synchronize()
    h.call("k2b_modified_beam_search_dev", raw, 1, B, T, K, tok, ts, n, sc, T)
    h.call("k2b_gather_results_nccl", tok, ts, n, sc, B, T, atok, ats, an, asc)
    h.sync()
    assert torch.equal(an, n) and torch.equal(asc, sc)
    for b in range(B):
        k = int(n[b])
        assert torch.equal(atok[b, :k], tok[b, :k]) and torch.equal(ats[b, :k], ts[b, :k])
    assert int(n.sum()) > 0
    # "async_gather": the gather of batch A runs on a side stream while batch B is searched INTO THE SAME result buffers; the
    # library orders B's back-trace (cluster engine) / B's whole call (other engines) behind the gather, so all_* hold A's results
    raw_b = torch.from_numpy(synth.make_frames(B, T, MID.encoder_dim, 6)).cuda()
    h.set_option("async_gather", 1)
    for prec in ("fp32", "bf16x3"):
        h.set_precision(_native.PREC_NAMES[prec])
        want = {}
        for name, x in (("a", raw), ("b", raw_b)):
            h.call("k2b_modified_beam_search_dev", x, 1, B, T, K, tok, ts, n, sc, T)
            h.sync()
            want[name] = (tok.clone(), ts.clone(), n.clone(), sc.clone())
        for rep in range(3):
            for o in (atok, ats, an, asc):
                o.fill_(-1)
            h.call("k2b_modified_beam_search_dev", raw, 1, B, T, K, tok, ts, n, sc, T)
            h.call("k2b_gather_results_nccl", tok, ts, n, sc, B, T, atok, ats, an, asc)
            h.call("k2b_modified_beam_search_dev", raw_b, 1, B, T, K, tok, ts, n, sc, T)
            if rep == 0:
                h.call("k2b_gather_join")
            h.sync()
            assert torch.equal(an, want["a"][2]) and torch.equal(asc, want["a"][3]), prec
            assert torch.equal(n, want["b"][2]) and torch.equal(sc, want["b"][3]), prec
            for b in range(B):
                k = int(an[b])
                assert torch.equal(atok[b, :k], want["a"][0][b, :k]) and torch.equal(ats[b, :k], want["a"][1][b, :k])
                k = int(n[b])
                assert torch.equal(tok[b, :k], want["b"][0][b, :k])
    h.close()


def test_device_call_overlaps_the_second_half_projection(built_lib):
    """k2b_modified_beam_search_dev on raw frames of a batch that fills the cluster kernel (B * K >= 512, T >= 64): the encoder_proj
    GEMM of the second half of the frames runs on a side stream under the search of the first half, the search is two launches
    with the state carried between them. Must equal the one-launch form (k2b_set_option("dev_chunks", 1)) bit for bit - also when
    calls follow each other without a host synchronisation, with ragged lengths, and on an odd frame count."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    B, K = 128, 4
    for T, lens in ((64, None), (77, None), (64, "ragged")):
        raws = [torch.from_numpy(synth.make_frames(B, T, MID.encoder_dim, 500 + i)).cuda() for i in range(2)]
        outs = {}
        for chunks in (1, 2):
            h.set_option("dev_chunks", chunks)
            res = []
            for i in range(2):                       # two calls back to back on the handle's stream
                tok = torch.zeros((B, T), dtype=torch.int64, device="cuda"); ts = torch.zeros((B, T), dtype=torch.int32, device="cuda")
                n = torch.zeros(B, dtype=torch.int32, device="cuda"); sc = torch.zeros(B, dtype=torch.float32, device="cuda")
                if lens is not None:
                    h.set_encoder_out_lens([(7 * b + 3 * i) % (T + 1) for b in range(B)])
                n0 = h.launch_count()
                h.call("k2b_modified_beam_search_dev", raws[i], 1, B, T, K, tok, ts, n, sc, T)
                res.append((tok, ts, n, sc, h.launch_count() - n0))
            h.sync()
            outs[chunks] = res
        for i in range(2):
            a, b = outs[1][i], outs[2][i]
            assert i == 0 or b[4] > a[4], "two projections + two search launches (the first call also packs the weights)"
            assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
            for s_ in range(B):
                k = int(a[2][s_])
                assert torch.equal(a[0][s_, :k], b[0][s_, :k]) and torch.equal(a[1][s_, :k], b[1][s_, :k])
        assert int(outs[2][0][2].sum()) > 0
    h.set_option("dev_chunks", -1)
    # against the oracle (8 streams of the last batch)
    T = 64
    raw = synth.make_frames(B, T, MID.encoder_dim, 777)
    d = torch.from_numpy(raw).cuda()
    tok = torch.zeros((B, T), dtype=torch.int64, device="cuda"); ts = torch.zeros((B, T), dtype=torch.int32, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda"); sc = torch.zeros(B, dtype=torch.float32, device="cuda")
    h.call("k2b_modified_beam_search_dev", d, 1, B, T, K, tok, ts, n, sc, T)
    h.sync()
    want = O.modified_beam_search(m, O.encoder_proj(m, raw[:8]), K)
    got_t = [tok[b, :int(n[b])].tolist() for b in range(8)]
    got_s = [ts[b, :int(n[b])].tolist() for b in range(8)]
    compare_streams(got_t, got_s, want, "device call, two chunks", allow_frac=0.25, scores=sc[:8].tolist(), T=T)
    h.close()


def test_backpointer_history_reconstructs_the_output(built_lib):
    """k2b_debug_backpointers: walking the history from the chosen hypothesis reproduces tokens / timestamps; no beam ever holds
    two hypotheses with the same token sequence (the hash dedupe checked against real sequences)."""
    m, w = model_and_weights(MID, blank_bias=0.99)
    h = make(MID, w)
    B, T, K = 16, 40, 4
    raw = synth.make_frames(B, T, MID.encoder_dim, 66)
    t, s, sc = h.modified_beam_search(raw, K)
    bp = h.debug_backpointers(B, T, K)
    for b in range(B):
        seqs = [((), ())]                                    # per slot of the previous frame: (tokens, timestamps)
        for fr in range(T):
            new = []
            for k in range(K):
                e = int(bp[b, fr, k])
                par, tok = (e >> 28) & 0xF, (e & 0x0FFFFFFF) - 1
                base = seqs[par]
                new.append((base[0] + ((tok,) if tok >= 0 else ()), base[1] + ((fr,) if tok >= 0 else ())))
            # a dead slot holds 0 = (parent 0, no token): it repeats a sequence that is already there (or stands behind all live
            # slots); only entries that are not of that form are hypotheses
            live = [q for k, q in enumerate(new) if int(bp[b, fr, k]) != 0 or k == 0]
            assert len({q[0] for q in live}) == len(live), f"stream {b} frame {fr}: two hypotheses share a token sequence"
            seqs = new
        assert (list(t[b]), list(s[b])) in [(list(q[0]), list(q[1])) for q in seqs]
    h.close()
