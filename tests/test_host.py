"""CPU: host-side logic of the reference-interface mirror (streams, chunking, list seeds, the fine-grained
loops) with the native handle replaced by an oracle-backed stub. Nothing here runs the CUDA library."""
import numpy as np
import pytest

from k2transducerasr_b200 import proj as P
from k2transducerasr_b200 import recognizer as R
from k2transducerasr_b200 import synth
from oracle import k2_oracle as O
from tests.helpers import SMALL, model_and_weights


class StubNative:
    """Serves the fine-grained entry points from the oracle (test double; never shipped)."""

    def __init__(self, model):
        self.m = model

    def decoder_proj(self, y, n=None):
        return O.decoder(self.m, None if y is None else np.asarray(y).reshape(-1, 2), n)

    def joiner_proj(self, enc, dec):
        return O.joiner(self.m, enc, dec)

    def encoder_proj(self, raw):
        return O.encoder_proj(self.m, raw)

    def close(self):
        pass


def make_offline(model, dims):
    p = P.OfflineProjOfB200.__new__(P.OfflineProjOfB200)
    p._dims, p._native, p._frames_are_raw, p._disposed = dims, StubNative(model), True, False
    p.Blank_id, p.Sos_eos_id, p.Unk_id = dims.blank_id, dims.sos_eos_id, dims.unk_id
    p.CustomMetadata = P.OfflineCustomMetadata(Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim)
    return p


def make_online(model, dims, chunk=8):
    p = P.OnlineProjOfB200.__new__(P.OnlineProjOfB200)
    p._dims, p._native, p._frames_are_raw, p._disposed = dims, StubNative(model), True, False
    p.Blank_id, p.Sos_eos_id, p.Unk_id = dims.blank_id, dims.sos_eos_id, dims.unk_id
    p.CustomMetadata = P.OnlineCustomMetadata(Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim, T=chunk)
    p.ChunkLength = p.ShiftLength = chunk
    p.FeatureDim, p.SampleRate = dims.encoder_dim, 16000
    p._layout, p._encoder_hook = None, None
    return p


@pytest.fixture(scope="module")
def setup():
    m, _ = model_and_weights(SMALL, blank_bias=0.6)
    raw = synth.make_frames(4, 24, SMALL.encoder_dim, 77)
    return m, raw, O.encoder_proj(m, raw)


def test_single_stream_fine_grained_loop_matches_oracle(setup):
    m, raw, enc = setup
    rec = R.OfflineRecognizer(make_offline(m, SMALL), fused=False)
    s = rec.CreateOfflineStream()
    assert s.Tokens == [0, 0]                                  # ref OfflineStream.cs:34
    s.AcceptFrames(raw[0, :10]); s.AcceptFrames(raw[0, 10:])
    rec.GetResult(s)
    want = O.greedy_search_single(m, enc[0])
    assert s.Tokens == want.tokens and s.Timestamps == want.timestamps and s.Tokens[:2] == [-1, 0]


@pytest.mark.parametrize("mode", ["compat", "per_stream"])
def test_batch_fine_grained_loop_matches_oracle(setup, mode):
    m, raw, enc = setup
    rec = R.OfflineRecognizer(make_offline(m, SMALL), fused=False, batch_mode=mode)
    streams = [rec.CreateOfflineStream() for _ in range(4)]
    for i, s in enumerate(streams):
        s.AcceptFrames(raw[i])
    rec.GetResults(streams)
    want = O.greedy_search_batch(m, enc, compat=(mode == "compat"))
    for s, w in zip(streams, want):
        assert s.Tokens == w.tokens and s.Timestamps == w.timestamps
        assert s.OfflineInputEntity.Speech is None             # RemoveSamples (ref :294)


def test_online_chunking_and_stream_removal(setup):
    m, raw, enc = setup
    rec = R.OnlineRecognizer(make_online(m, SMALL, 8), fused=False)
    streams = [rec.CreateOnlineStream() for _ in range(3)]
    streams[0].AcceptFrames(raw[0]); streams[1].AcceptFrames(raw[1]); streams[2].AcceptFrames(raw[2, :5])
    lst = list(streams)
    rec.GetResults(lst)
    assert len(lst) == 2 and streams[2] not in lst             # ref OnlineRecognizer.cs:117-120
    for _ in range(2):
        rec.GetResults(lst)
    hyps, toks, tss = [[0, 0]] * 2, [[0, 0]] * 2, [[], []]
    for c in range(3):
        res = O.greedy_search_online_chunk(m, enc[:2, 8 * c:8 * c + 8], hyps, toks)
        hyps, toks = [r.hyp for r in res], [r.tokens for r in res]
        for b in range(2):
            tss[b] += res[b].timestamps
    for b in range(2):
        assert streams[b].Tokens == toks[b] and streams[b].Timestamps == tss[b]
        assert streams[b].Hyp.tolist() == hyps[b]
        assert all(0 <= t < 8 for t in streams[b].Timestamps)  # chunk-local (ref :184)
    assert streams[0].GetDecodeChunk() is None


def test_text_decoding_skips_specials():
    syms = ["<blk> 0", "<sos/eos> 1", "<unk> 2", "▁HE 3", "LLO 4", "▁WORLD 5"]
    text, toks = R._decode_text([-1, 0, 3, 4, 2, 5], syms)
    # ref OfflineRecognizer.cs:443-446 stops at the first id 2; CheckText strips the spaces of tag-free text (ref :497-499)
    assert text == "hello" and toks == ["▁HE", "LLO"]


def test_pad_batch_pads_with_zero_frames():
    a = P.OfflineInputEntity(Speech=np.ones((3, 4), np.float32))
    b = P.OfflineInputEntity(Speech=np.ones((5, 4), np.float32))
    x = P._pad_batch([a, b], 4)
    assert x.shape == (2, 5, 4) and x[0, 3:].sum() == 0


def test_bench_clock_sampler_summary():
    """bench.py's clock / throttle-reason summary (rows as nvidia-smi's csv or the NVML sampler produce them)."""
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location("bench_mod", pathlib.Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.ClockSampler(0)
    c.rows = [["1965", "1965", "700.0", "Not Active", "Not Active", "Not Active", "Not Active"],
              ["1950", "1965", "990.0", "Not Active", "Not Active", "Not Active", "Active"],
              ["1965", "1965", "710.0", "Not Active", "Not Active", "Not Active", "Not Active"]]
    s = c.summary()
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["samples"] == 3
    assert bench.ClockSampler(0).summary()["sm_mhz"] is None
