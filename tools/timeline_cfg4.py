"""Per-SM clock64 timeline of the per-frame beam path on cfg4 (joiner / fused merge kernels), run on the GPU box: where a frame
step's time goes, including the gaps between the kernels (which CUDA events and ncu cannot see)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build  # noqa: E402

build.build()
cfg = synth.CONFIGS["cfg4"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
T = 40
raw = synth.make_frames(cfg.streams, T, d.encoder_dim, cfg.seed)
h.modified_beam_search(raw, 4, enc_is_raw=True)
h.debug_timeline()
h.modified_beam_search(raw, 4, enc_is_raw=True)
tl = h.debug_timeline()[:T]                      # [frame][sm][8]
BIG = 0x7fffffffffffffff
rows = []
for t in range(5, T - 2):
    for sm in range(148):
        a = tl[t, sm]; b = tl[t + 1, sm]
        if a[0] == 0 or a[3] == 0 or a[4] == BIG or a[7] == 0 or b[0] == 0:
            continue
        rows.append([a[1] - a[0], a[2] - a[0], a[3] - a[0], a[4] - a[3], a[5] - a[3], a[6] - a[5], a[7] - a[6], b[0] - a[7],
                     b[0] - a[0], b[2] - a[2]])
r = np.array(rows, dtype=np.float64)
names = ["joiner start -> wait done", "joiner start -> first accumulator", "joiner start -> end (epilogue thread)",
         "joiner end -> merge first CTA start", "joiner end -> merge wait done", "merge wait done -> merge done (last CTA)",
         "merge done -> operand written (last CTA)", "merge end -> next joiner start", "frame period (joiner start to start)",
         "frame period (first accumulator to first accumulator)"]
print(f"{len(rows)} (frame, SM) samples; cycles: median / p10 / p90")
for i, n in enumerate(names):
    print(f"  {n:55s} {np.median(r[:, i]):9.0f} {np.percentile(r[:, i], 10):9.0f} {np.percentile(r[:, i], 90):9.0f}")
