"""Two cfg2 modified_beam_search calls on projected frames, one cluster-kernel launch of 250 frames each (pipe_chunks = 1: the
host-pointer call is not cut into time chunks). Used under ncu."""
import sys
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build

build.build()
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES[prec])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
h.set_option("pipe_chunks", 1)
raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
enc = h.encoder_proj(raw)
for _ in range(2):
    toks, tss, sc = h.modified_beam_search(enc, cfg.beam, enc_is_raw=False)
print("ok", sum(len(t) for t in toks), h.launch_count())
h.close()
