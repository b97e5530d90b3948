"""cfg3 (512 online streams, vocab 2000, chunks of 8 frames): device time of the cluster kernel per chunk call vs the whole call."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
cfg = synth.CONFIGS["cfg3"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
B, Tc = cfg.streams, cfg.frames
raw = synth.make_frames(B, Tc * 4, d.encoder_dim, cfg.seed)
hyp = np.zeros((B, 2), np.int64)
for c in range(2):
    t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
h.profile_enable(True)
t0 = time.perf_counter()
for c in range(2, 4):
    t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, Tc * c:Tc * c + Tc]), hyp, enc_is_raw=True)
dt = (time.perf_counter() - t0) / 2
nl, ms = h.profile_read()
print(f"cluster kernel: {nl} launches, {1e3 * ms / nl:.1f} us per chunk launch; host call {1e6 * dt:.0f} us (incl. H2D/D2H)")
for T in (1, 8, 32):
    e = np.ascontiguousarray(np.tile(raw[:, :1], (1, T, 1)))
    hyp2 = np.zeros((B, 2), np.int64)
    h.greedy_online_chunk(e, hyp2, enc_is_raw=True)
    h.profile_read()
    h.greedy_online_chunk(e, hyp2, enc_is_raw=True)
    nl, ms = h.profile_read()
    print(f"T'={T}: {1e3 * ms / nl:.1f} us per launch")
h.close()
