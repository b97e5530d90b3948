"""cfg4 with the per-frame launches (K2B_NO_MEGA=1) for ncu captures of the small kernels: operand gather, merge, back-trace."""
import os
import sys
os.environ["K2B_NO_MEGA"] = "1"
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
cfg = synth.CONFIGS["cfg4"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raw = synth.make_frames(cfg.streams, 6, d.encoder_dim, cfg.seed)
h.modified_beam_search(raw, 4, enc_is_raw=True)
h.modified_beam_search(raw, 4, enc_is_raw=True)
print("done")
