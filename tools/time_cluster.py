"""Device time of the cluster kernel alone on cfg2 (projected frames resident, k2b_profile_* CUDA events on the launch stream)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
for prec in (sys.argv[1:] or ["bf16x3"]):
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_NAMES[prec])
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
    enc = h.encoder_proj(raw)
    for _ in range(3):
        h.modified_beam_search(enc, cfg.beam, enc_is_raw=False)
    h.profile_enable(True)
    n = 10
    for _ in range(n):
        toks, tss, sc = h.modified_beam_search(enc, cfg.beam, enc_is_raw=False)
    nl, ms = h.profile_read()
    h.profile_enable(False)
    us = ms * 1e3 / nl
    print(f"{prec}: {nl} launches, {us:.1f} us per launch, {us / cfg.frames:.2f} us per frame step, "
          f"{cfg.streams * cfg.frames / us:.2f} M frames/s (kernel only); tokens {sum(len(t) for t in toks)}")
    h.close()
