"""The persistent large-vocabulary kernel with its records announced by counters (tagged_records = 0) or by their own epoch tags
(1, the default): device time of the whole search and equality of the outputs. cfg4 (beam 4, V = 5537) and cfg3's shape as one 64-frame search
(greedy, V = 2000, 512 streams)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
for name, T, K in (("cfg4", 250, 4), ("cfg3", 64, 1)):
    cfg = synth.CONFIGS[name]; d = cfg.dims
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_BF16X3)
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    h.set_option("pipe_chunks", 1)
    raw = synth.make_frames(cfg.streams, T, d.encoder_dim, cfg.seed)
    enc = h.encoder_proj(raw)
    outs = {}
    for tagged in (0, 1, 0, 1):
        h.set_option("tagged_records", tagged)
        run = (lambda: h.modified_beam_search(enc, K, enc_is_raw=False)) if K > 1 else \
              (lambda: h.greedy_offline(enc, _native.GREEDY_PER_STREAM, enc_is_raw=False))
        run()
        h.profile_enable(True)
        for _ in range(3):
            out = run()
        n, ms = h.profile_read()
        h.profile_enable(False)
        outs[tagged] = out
        print(f"{name} tagged_records={tagged}: {1e3 * ms / max(n, 1):8.1f} us per launch ({1e3 * ms / max(n, 1) / T:.2f} us per frame)", flush=True)
    same = outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    print(f"{name}: outputs identical between the two forms: {same}")
    h.close()
