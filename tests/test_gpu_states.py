"""GPU (-m gpu): on-device streaming state - k2b_stack_states / k2b_unstack_states against the oracle's restatement of the
reference's Array.Copy loops (bit-exact: it is a copy), on a zipformer2-shaped cache layout and on awkward shapes."""
import numpy as np
import pytest
import torch

from k2transducerasr_b200 import _native
from oracle import k2_oracle as O

pytestmark = pytest.mark.gpu


def zipformer2_layout():
    """Cache tensors of one stream for the streaming zipformer2 of README.EN.md:14 (metadata keys of ref
    Model/OnlineCustomMetadata.cs): per layer key / nonlin_attn / val1 / val2 / conv1 / conv2, then embed_states and
    processed_lens. Returns (item_len, axis_len of stack_states)."""
    layers, dims = (2, 2, 3, 4, 3, 2), (192, 256, 384, 512, 384, 256)
    heads, qd, vd = (4, 4, 4, 8, 4, 4), (32,) * 6, (12,) * 6
    left, kern = (64, 32, 16, 8, 16, 32), (31, 31, 15, 15, 15, 31)
    il, ax = [], []
    for i, nl in enumerate(layers):
        key, val, pad = qd[i] * heads[i], vd[i] * heads[i], kern[i] // 2
        for _ in range(nl):
            il += [left[i] * key, left[i] * (3 * dims[i] // 4), left[i] * val, left[i] * val, dims[i] * pad, dims[i] * pad]
            ax += [key, left[i], val, val, dims[i] * pad, dims[i] * pad]           # ref OnlineProjOfZipformer2.cs:236-305
    il += [128 * 3 * 19, 1]
    ax += [128 * 3 * 19, 1]                                                          # ref :322, :341
    return il, ax


@pytest.mark.parametrize("layout", ["zipformer2", "odd"])
def test_stack_unstack_states_bit_exact(built_lib, layout):
    if layout == "zipformer2":
        il, ax = zipformer2_layout()
        B, pool = 5, 9
    else:
        il, ax = [6, 35, 1, 4096 + 8, 3 * 7], [3, 5, 1, 8, 7]         # unaligned offsets, scalar path, more than one chunk
        B, pool = 7, 7
    rng = np.random.default_rng(3)
    h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
    h.state_pool_create(il, pool)
    states = [[rng.standard_normal(n).astype(np.float32) for n in il] for _ in range(pool)]
    for s in range(pool):
        h.state_pool_put(s, np.concatenate(states[s]))
    np.testing.assert_array_equal(h.state_pool_get(2), np.concatenate(states[2]))
    slots = list(rng.permutation(pool)[:B])
    n = h.state_pool_stacked_floats(B)
    buf = torch.zeros(n, dtype=torch.float32, device="cuda")
    h.stack_states_dev(slots, ax, buf.data_ptr())
    h.sync()
    got = buf.cpu().numpy()
    want = O.stack_states([states[s] for s in slots], ax)
    off = 0
    for i, L in enumerate(il):
        np.testing.assert_array_equal(got[B * off:B * off + B * L], want[i], err_msg=f"tensor {i}")
        off += (L + 3) & ~3
    # unstack with the reference's own axis of that direction (cached_nonlin_attn differs: ref :409) into fresh slots
    ax_un = list(ax)
    new = rng.standard_normal(n).astype(np.float32)
    newbuf = torch.from_numpy(new).cuda()
    h.unstack_states_dev(slots, ax_un, newbuf.data_ptr())
    h.sync()
    stacked_list, off = [], 0
    for L in il:
        stacked_list.append(new[B * off:B * off + B * L])
        off += (L + 3) & ~3
    items = O.unstack_states(stacked_list, B, ax_un)
    for k, s in enumerate(slots):
        np.testing.assert_array_equal(h.state_pool_get(s), np.concatenate(items[k]), err_msg=f"slot {s}")
    untouched = [s for s in range(pool) if s not in slots]
    for s in untouched:
        np.testing.assert_array_equal(h.state_pool_get(s), np.concatenate(states[s]))
    with pytest.raises(_native.K2bError):
        h.unstack_states_dev([slots[0], slots[0]], ax, newbuf.data_ptr())          # a slot twice
    with pytest.raises(_native.K2bError):
        h.stack_states_dev([pool], ax, buf.data_ptr())                              # slot out of range
    h.close()


def test_stack_states_bandwidth_cfg3(built_lib):
    """cfg3 scale: 512 concurrent streams. Round trip is exact; the copy's HBM rate is printed (2 * B * Ls * 4 bytes per call)."""
    il, ax = zipformer2_layout()
    B = 512
    h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
    h.state_pool_create(il, B)
    n = h.state_pool_stacked_floats(B)
    a = torch.randn(n, dtype=torch.float32, device="cuda")
    b = torch.zeros_like(a)
    slots = list(range(B))
    h.unstack_states_dev(slots, ax, a.data_ptr())
    h.stack_states_dev(slots, ax, b.data_ptr())
    h.sync()
    # padding floats between tensors are not part of the tensors: compare tensor by tensor
    off = 0
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    for L in il:
        np.testing.assert_array_equal(an[B * off:B * off + B * L], bn[B * off:B * off + B * L])
        off += (L + 3) & ~3
    stream = torch.cuda.Stream()                      # CUDA events only see the stream they are recorded on
    h.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        h.stack_states_dev(slots, ax, b.data_ptr())
    e0.record(stream)
    for _ in range(10):
        h.stack_states_dev(slots, ax, b.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = 2 * B * sum(il) * 4 / (ms * 1e-3) / 1e9
    print(f"stack_states B={B}: {sum(il) * 4 / 1e6:.2f} MB per stream, {ms * 1e3:.1f} us per call, {gbs:.0f} GB/s")
    assert gbs > 1000
    h.close()
