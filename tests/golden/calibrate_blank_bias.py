"""Find the joiner blank bias that makes ~75 % of greedy frames blank for each config's dims.
Uses the CPU oracle (test infrastructure); the constants it prints are pasted into
k2transducerasr_b200/synth.py::CONFIGS.  Run from the repo root: python tests/golden/calibrate_blank_bias.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import synth
from oracle import k2_oracle as O


def blank_fraction(dims, bias, B=6, T=80):
    w = synth.make_weights(dims, blank_bias=bias)
    m = O.Model.from_dict(w)
    if dims.encoder_dim:
        enc = O.encoder_proj(m, synth.make_frames(B, T, dims.encoder_dim, 11))
    else:
        enc = synth.make_frames(B, T, dims.joiner_dim, 11)
    r = O.greedy_search_batch(m, enc, compat=False)
    return 1.0 - np.mean([len(x.appended) for x in r]) / T


if __name__ == "__main__":
    for name in ("cfg1", "cfg2", "cfg3", "cfg4"):
        dims = synth.CONFIGS[name].dims
        lo, hi = 0.0, 3.0
        for _ in range(9):
            mid = 0.5 * (lo + hi)
            if blank_fraction(dims, mid) < 0.75:
                lo = mid
            else:
                hi = mid
        b = round(0.5 * (lo + hi), 2)
        print(name, "blank_bias", b, "blank fraction", round(blank_fraction(dims, b), 3))
    # CTC: direct on the log-prob generator
    for bias in (3.0, 4.0, 5.0, 6.0, 7.0):
        lp = synth.make_ctc_logp(8, 100, 2000, 5, blank_bias=bias)
        print("ctc bias", bias, "blank fraction", float((lp.argmax(-1) == 0).mean()))
