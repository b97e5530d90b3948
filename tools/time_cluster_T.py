"""Cluster-kernel time as a function of the frame count (fixed launch cost vs per-frame cost) on cfg2's shape."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
cfg = synth.CONFIGS["cfg2"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
enc = h.encoder_proj(raw)
res = {}
for T in (2, 10, 50, 250):
    e = np.ascontiguousarray(enc[:, :T])
    for _ in range(3):
        h.modified_beam_search(e, cfg.beam, enc_is_raw=False)
    h.profile_enable(True)
    for _ in range(10):
        h.modified_beam_search(e, cfg.beam, enc_is_raw=False)
    nl, ms = h.profile_read()
    h.profile_enable(False)
    res[T] = ms * 1e3 / nl
    print(f"T={T:4d}: {res[T]:8.1f} us per launch")
b = (res[250] - res[50]) / 200
print(f"per-frame {b:.3f} us ({b * 1965:.0f} cycles at 1.965 GHz), fixed {res[250] - 250 * b:.1f} us")
for B in (8, 32, 128):
    e = np.ascontiguousarray(enc[:B])
    for _ in range(3):
        h.modified_beam_search(e, cfg.beam, enc_is_raw=False)
    h.profile_enable(True)
    for _ in range(10):
        h.modified_beam_search(e, cfg.beam, enc_is_raw=False)
    nl, ms = h.profile_read()
    h.profile_enable(False)
    print(f"B={B:4d} streams ({B // 8} clusters): {ms * 1e3 / nl:8.1f} us per launch, {ms * 1e3 / nl / cfg.frames:.2f} us per frame")
h.close()
