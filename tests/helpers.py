"""Shared checkers of the parity tests. The oracle (oracle/k2_oracle.py) is the reference side."""
from __future__ import annotations

import numpy as np

from k2transducerasr_b200 import synth
from oracle import k2_oracle as O

NEAR_TIE = 1e-4      # BASELINE.json north_star: frames with a top-2 gap under 1e-4 are listed separately
SCORE_TOL = 1e-3     # hypothesis log-probs: 1e-3 absolute in fp32

SMALL = synth.ModelDims(vocab_size=53, joiner_dim=32, decoder_dim=32, encoder_dim=48)
MID = synth.ModelDims(vocab_size=500, joiner_dim=512, decoder_dim=512, encoder_dim=256)


def model_and_weights(dims, blank_bias=0.0, seed=synth.WEIGHT_SEED, **kw):
    w = synth.make_weights(dims, seed=seed, blank_bias=blank_bias)
    return O.Model.from_dict(w, blank_id=dims.blank_id, sos_eos_id=dims.sos_eos_id, unk_id=dims.unk_id,
                             context_size=dims.context_size, **kw), w


def compare_streams(got_tokens, got_ts, want, what="", allow_frac=0.02):
    """Token / timestamp sequences must be identical per stream; a stream may differ only if the oracle met a
    near-tie decision (gap < NEAR_TIE) while decoding it. Returns the list of excused streams."""
    assert len(got_tokens) == len(want)
    excused = []
    for b, r in enumerate(want):
        n = len(r.appended)
        want_ts = list(r.timestamps)[len(r.timestamps) - n:] if n else []
        if list(got_tokens[b]) == list(r.appended) and list(got_ts[b]) == want_ts:
            continue
        assert r.min_gap < NEAR_TIE, (
            f"{what}: stream {b} differs from the oracle without a near tie (min gap {r.min_gap:.3g}):\n"
            f" got  {list(got_tokens[b])[:40]}\n want {list(r.appended)[:40]}")
        excused.append(b)
    assert len(excused) <= max(1, int(allow_frac * len(want))), f"{what}: too many near-tie streams: {excused}"
    return excused
