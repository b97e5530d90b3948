"""Per-phase cycle breakdown of the 16-CTA cluster kernel on cfg3 (online greedy, vocab 2000), CTA 0 of cluster 0."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
from tools.phase_profile import names
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
B, T = cfg.streams, 32
raw = synth.make_frames(B, T, d.encoder_dim, cfg.seed)
hyp = np.zeros((B, 2), np.int64)
h.greedy_online_chunk(raw, hyp, enc_is_raw=True)
h.cluster_phase_cycles()
h.greedy_online_chunk(raw, hyp, enc_is_raw=True)
cyc = h.cluster_phase_cycles()
tot = cyc[:8].sum() + cyc[15:19].sum()
print(f"== {cfg.name}: {tot / T:.0f} cycles per frame step (CTA 0)")
for n, c in zip(names, cyc[:19]):
    print(f"   {n:44s} {c / T:8.0f} cyc  {100.0 * c / tot:5.1f} %")
h.close()
