"""Small tensor-path workload for compute-sanitizer memcheck (one tool per call, smallest case)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth
for dims, B, T in ((synth.ModelDims(97, 64, 48, 64), 3, 4), (synth.ModelDims(500, 512, 512, 256), 5, 3)):
    h = _native.Handle(vocab_size=dims.vocab_size, joiner_dim=dims.joiner_dim, decoder_dim=dims.decoder_dim,
                       encoder_dim=dims.encoder_dim, precision=_native.PREC_BF16X3)
    h.load_weights(synth.make_weights(dims, blank_bias=0.5))
    raw = synth.make_frames(B, T, dims.encoder_dim, 3)
    print(dims.vocab_size, h.modified_beam_search(raw, 4, enc_is_raw=True)[0])
    print(h.greedy_offline(raw, _native.GREEDY_PER_STREAM, enc_is_raw=True)[0])
    print(h.greedy_offline(raw, _native.GREEDY_BATCH_COMPAT, enc_is_raw=True)[0])
    print(h.modified_beam_search(raw, 3, enc_is_raw=True)[0])
    lp = synth.make_ctc_logp(3, 40, 101, 5)
    print(h.ctc_greedy(lp)[0][0][:5])
    h.close()
print("done")
