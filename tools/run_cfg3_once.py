"""Two online greedy chunks of cfg3's shape (512 streams, V = 2000, 8 frames) for ncu captures of the persistent greedy kernel."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
cfg = synth.CONFIGS["cfg3"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raw = synth.make_frames(cfg.streams, 2 * cfg.frames, d.encoder_dim, cfg.seed)
hyp = np.zeros((cfg.streams, 2), np.int64)
for c in range(2):
    t, s, hyp = h.greedy_online_chunk(np.ascontiguousarray(raw[:, c * cfg.frames:(c + 1) * cfg.frames]), hyp, enc_is_raw=True)
print("done", sum(len(x) for x in t))
