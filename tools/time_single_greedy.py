"""cfg1 (one stream, 250 frames, V = 500) on the register-resident greedy kernel: device time per launch, the tcgen05 cluster kernel
as beam 1 beside it, and the phase cycle counts the kernel prints (k2b_set_option("single_greedy", 2))."""
import sys
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
cfg = synth.CONFIGS["cfg1"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_BF16X3)
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raw = synth.make_frames(1, cfg.frames, d.encoder_dim, cfg.seed)
enc = h.encoder_proj(raw)
base = None
for opt, name in ((0, "cluster kernel, beam 1"), (1, "register kernel"), (2, "register kernel + phase print")):
    h.set_option("single_greedy", opt)
    out = h.greedy_offline(enc, _native.GREEDY_SINGLE, enc_is_raw=False)
    h.profile_enable(True)
    for _ in range(5 if opt < 2 else 1):
        out = h.greedy_offline(enc, _native.GREEDY_SINGLE, enc_is_raw=False)
    n, ms = h.profile_read()
    h.profile_enable(False)
    base = base or out
    print(f"{name}: {1e3 * ms / max(n, 1):7.1f} us per launch, {1e3 * ms / max(n, 1) / cfg.frames:.2f} us per frame; same symbols as the "
          f"cluster kernel: {out[0] == base[0]}", flush=True)
h.close()
