"""One warm-up and one timed modified_beam_search over cfg4 (for ncu captures): python tools/run_cfg4_once.py [precision] [frames]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = synth.CONFIGS["cfg4"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES[prec])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raw = synth.make_frames(cfg.streams, T, d.encoder_dim, cfg.seed)
h.modified_beam_search(raw, 4, enc_is_raw=True)
h.profile_enable(True)
h.modified_beam_search(raw, 4, enc_is_raw=True)
n, ms = h.profile_read()
print(f"{prec}: {n} bracketed launches, avg {1e3 * ms / max(n, 1):.1f} us ({1e3 * ms / max(n, 1) / T:.2f} us per frame)")
