"""clock64 timeline of CTA 0 of the persistent beam-search kernel on cfg4 (run on the GPU box): per frame, when each role got
what it waited for. Columns are cycles relative to the frame's first stamp."""
import sys

import numpy as np

sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build  # noqa: E402

build.build()
cfg = synth.CONFIGS["cfg4"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES[sys.argv[1] if len(sys.argv) > 1 else "bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
T = 40
raw = synth.make_frames(cfg.streams, T, d.encoder_dim, cfg.seed)
h.modified_beam_search(raw, 4, enc_is_raw=True)
h.debug_timeline()
h.modified_beam_search(raw, 4, enc_is_raw=True)
raw_tl = h.debug_timeline().reshape(-1)
tl = raw_tl[:T * 16].reshape(T, 16)
mg = raw_tl[640:640 + T * 8].reshape(T, 8)
names = ["ldA", "accA", "epiA", "ldB", "accB", "epiB", "m1go", "m1end", "m2go", "m2end", "A:pass1", "A:pass2", "A:comb", "-", "m1:B", "m1:C"]
print("frame  period " + " ".join(f"{n:>7s}" for n in names))
for t in range(8, 16):
    base = tl[t, 1]
    print(f"{t:5d} {tl[t + 1, 1] - tl[t, 1]:7d} " + " ".join(f"{tl[t, i] - base:7d}" for i in range(len(names))))
per = np.diff(tl[5:T - 1, 1])
print("median frame period (cycles):", np.median(per))

print("merge of the first stream, cycles after its go: lse, cands, A done, B rounds, extension, dedupe+logadd, B done(6), C done(7)")
for t in range(8, 16):
    print(f"{t:5d} " + " ".join(f"{mg[t, i] - tl[t, 6]:7d}" for i in range(8)))
