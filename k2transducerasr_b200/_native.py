"""ctypes binding of libk2b200.so — the Python stand-in for the C# P/Invoke layer (`NativeMethods`, see
INTEGRATION.md). It declares exactly the entry points of include/k2b200.h and raises if the shared
library is missing: there is no CPU fallback and no path around the CUDA library."""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libk2b200.so"
HEADER_PATH = PKG.parent / "include" / "k2b200.h"

K2B_OK, K2B_ERR_INVALID, K2B_ERR_CUDA, K2B_ERR_STATE, K2B_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
GREEDY_SINGLE, GREEDY_BATCH_COMPAT, GREEDY_PER_STREAM = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
NEGID_MASK, NEGID_WRAP = 0, 1
PREC_NAMES = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}


class K2bError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libk2b200 status {status}: {message}")
        self.status = status
        self.message = message


class K2bConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "device", "vocab_size", "joiner_dim", "decoder_dim", "encoder_dim", "context_size",
        "blank_id", "sos_eos_id", "unk_id", "max_streams", "max_frames", "max_beam", "neg_id_mode",
        "precision", "reserved")]


_P = C.c_void_p
_I = C.c_int32
_SIGS = {
    "k2b_abi_version": (C.c_int32, []),
    "k2b_create": (C.c_int32, [C.POINTER(K2bConfig), C.POINTER(_P)]),
    "k2b_destroy": (C.c_int32, [_P]),
    "k2b_last_error": (C.c_char_p, [_P]),
    "k2b_load_weights": (C.c_int32, [_P] + [_P] * 8),
    "k2b_set_precision": (C.c_int32, [_P, _I]),
    "k2b_set_stream": (C.c_int32, [_P, _P]),
    "k2b_sync": (C.c_int32, [_P]),
    "k2b_launch_count": (C.c_int64, [_P]),
    "k2b_reset_launch_count": (C.c_int32, [_P]),
    "k2b_profile_enable": (C.c_int32, [_P, _I]),
    "k2b_profile_read": (C.c_int32, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "k2b_decoder_proj": (C.c_int32, [_P, _P, _I, _P]),
    "k2b_decoder_proj_dev": (C.c_int32, [_P, _P, _I, _P]),
    "k2b_joiner_proj": (C.c_int32, [_P, _P, _P, _I, _P]),
    "k2b_joiner_proj_dev": (C.c_int32, [_P, _P, _P, _I, _P]),
    "k2b_encoder_proj": (C.c_int32, [_P, _P, _I, _P]),
    "k2b_encoder_proj_dev": (C.c_int32, [_P, _P, _I, _P]),
    "k2b_set_encoder_out_lens": (C.c_int32, [_P, _P, _I]),
    "k2b_greedy_offline": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I]),
    "k2b_greedy_offline_dev": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I]),
    "k2b_greedy_online_chunk": (C.c_int32, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _I]),
    "k2b_greedy_online_chunk_dev": (C.c_int32, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _I]),
    "k2b_modified_beam_search": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I]),
    "k2b_modified_beam_search_dev": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I]),
    "k2b_ctc_greedy": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I]),
    "k2b_ctc_greedy_dev": (C.c_int32, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I]),
    "k2b_set_option": (C.c_int32, [_P, C.c_char_p, _I]),
    "k2b_get_stat": (C.c_int32, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "k2b_host_alloc": (C.c_int32, [C.POINTER(_P), C.c_int64]),
    "k2b_host_free": (C.c_int32, [_P]),
    "k2b_host_register": (C.c_int32, [_P, C.c_int64]),
    "k2b_host_unregister": (C.c_int32, [_P]),
    "k2b_set_context_graph": (C.c_int32, [_P, _P, _P, _P, _I]),
    "k2b_beam_pool_create": (C.c_int32, [_P, _I, _I, _I]),
    "k2b_beam_pool_reset": (C.c_int32, [_P, _I, _P]),
    "k2b_modified_beam_search_online_chunk": (C.c_int32, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I]),
    "k2b_modified_beam_search_online_chunk_dev": (C.c_int32, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I]),
    "k2b_nccl_unique_id": (C.c_int32, [_P]),
    "k2b_nccl_init": (C.c_int32, [_P, _P, _I, _I]),
    "k2b_gather_results_nccl": (C.c_int32, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "k2b_gather_join": (C.c_int32, [_P]),
    "k2b_debug_backpointers": (C.c_int32, [_P, _P, _I, _I, _I]),
    "k2b_selftest_umma": (C.c_int32, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "k2b_state_pool_create": (C.c_int32, [_P, _P, _I, _I]),
    "k2b_state_pool_stacked_floats": (C.c_int64, [_P, _I]),
    "k2b_state_pool_put": (C.c_int32, [_P, _I, _P]),
    "k2b_state_pool_get": (C.c_int32, [_P, _I, _P]),
    "k2b_stack_states": (C.c_int32, [_P, _P, _I, _P, _P]),
    "k2b_unstack_states": (C.c_int32, [_P, _P, _I, _P, _P]),
    "k2b_selftest_umma2": (C.c_int32, [_P, _P, _P, _I, _I, _I, _P]),
    "k2b_cluster_phase_cycles": (C.c_int32, [_P, _P]),
    "k2b_debug_timeline": (C.c_int32, [_P, _P]),
    "k2b_selftest_umma_bench": (C.c_int32, [_P, _I, _I, _I, _P]),
    "k2b_selftest_collectives": (C.c_int32, [_P, _P]),
    "k2b_selftest_dsmem_bw": (C.c_int32, [_P, _I, _I, _I, _P]),
    "k2b_selftest_cluster": (C.c_int32, [_P, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
}


def header_symbols() -> list[str]:
    """Every function include/k2b200.h declares (used by the ABI test)."""
    text = HEADER_PATH.read_text()
    return sorted(set(re.findall(r"K2B_API\s+[\w\s\*]+?\b(k2b_\w+)\s*\(", text)))


_lib = None
_HOST_ALLOCS: dict = {}


def lib() -> C.CDLL:
    """Load libk2b200.so (built in-tree by k2transducerasr_b200.build). Raises when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -m k2transducerasr_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    handle = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def _ptr(a) -> Optional[int]:
    """Host numpy array -> address; int -> device/host address as is; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host memory)
        return int(a.data_ptr())
    raise TypeError(type(a))


class Handle:
    """Owns one k2b_handle (one device, one stream; not thread-safe)."""

    def __init__(self, *, vocab_size: int, joiner_dim: int = 512, decoder_dim: int = 512, encoder_dim: int = 0,
                 context_size: int = 2, blank_id: int = 0, sos_eos_id: int = 1, unk_id: int = 2, device: int = 0,
                 max_streams: int = 0, max_frames: int = 0, max_beam: int = 4, neg_id_mode: int = NEGID_MASK,
                 precision: int = PREC_FP32):
        self._lib = lib()
        self.cfg = K2bConfig(C.sizeof(K2bConfig), device, vocab_size, joiner_dim, decoder_dim, encoder_dim,
                             context_size, blank_id, sos_eos_id, unk_id, max_streams, max_frames, max_beam,
                             neg_id_mode, precision, 0)
        self._h = _P()
        st = self._lib.k2b_create(C.byref(self.cfg), C.byref(self._h))
        if st != K2B_OK:
            msg = self._lib.k2b_last_error(None)
            self._h = _P()
            raise K2bError(st, msg.decode() if msg else "k2b_create failed")
        self._keep = []

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, st: int):
        if st != K2B_OK:
            msg = self._lib.k2b_last_error(self._h)
            raise K2bError(st, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.k2b_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def V(self): return self.cfg.vocab_size
    @property
    def J(self): return self.cfg.joiner_dim
    @property
    def E(self): return self.cfg.encoder_dim

    def load_weights(self, w: dict):
        def f(name):
            a = w.get(name)
            return None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        arrs = [f(k) for k in ("emb", "conv_w", "dec_proj_w", "dec_proj_b", "enc_proj_w", "enc_proj_b", "out_w", "out_b")]
        self._check(self._lib.k2b_load_weights(self._h, *[_ptr(a) for a in arrs]))

    def set_precision(self, precision):
        if isinstance(precision, str):
            precision = PREC_NAMES[precision]
        self._check(self._lib.k2b_set_precision(self._h, precision))

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self._lib.k2b_set_stream(self._h, cuda_stream))

    def sync(self):
        self._check(self._lib.k2b_sync(self._h))

    def set_option(self, name: str, value: int):
        self._check(self._lib.k2b_set_option(self._h, name.encode(), int(value)))

    def get_stat(self, name: str) -> float:
        v = C.c_double(0.0)
        self._check(self._lib.k2b_get_stat(self._h, name.encode(), C.byref(v)))
        return float(v.value)

    def launch_count(self) -> int:
        return int(self._lib.k2b_launch_count(self._h))

    def reset_launch_count(self):
        self._check(self._lib.k2b_reset_launch_count(self._h))

    def profile_enable(self, on: bool):
        self._check(self._lib.k2b_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        n, ms = C.c_int64(0), C.c_double(0.0)
        self._check(self._lib.k2b_profile_read(self._h, C.byref(n), C.byref(ms)))
        return int(n.value), float(ms.value)

    # -- fine-grained (host numpy in/out) ---------------------------------------------------------------
    def decoder_proj(self, y: Optional[np.ndarray], n: Optional[int] = None) -> np.ndarray:
        if y is not None:
            y = np.ascontiguousarray(y, dtype=np.int64).reshape(-1, self.cfg.context_size)
            n = y.shape[0]
        out = np.empty((n, self.J), np.float32)
        self._check(self._lib.k2b_decoder_proj(self._h, _ptr(y), n, _ptr(out)))
        return out

    def joiner_proj(self, enc: np.ndarray, dec: np.ndarray) -> np.ndarray:
        enc = np.ascontiguousarray(enc, dtype=np.float32).reshape(-1, self.J)
        dec = np.ascontiguousarray(dec, dtype=np.float32).reshape(-1, self.J)
        if enc.shape != dec.shape:
            raise ValueError("encoder_out and decoder_out must have the same number of rows")
        n = enc.shape[0]
        out = np.empty((n, self.V), np.float32)
        self._check(self._lib.k2b_joiner_proj(self._h, _ptr(enc), _ptr(dec), n, _ptr(out)))
        return out

    def encoder_proj(self, raw: np.ndarray) -> np.ndarray:
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        lead = raw.shape[:-1]
        flat = raw.reshape(-1, self.E)
        out = np.empty((flat.shape[0], self.J), np.float32)
        self._check(self._lib.k2b_encoder_proj(self._h, _ptr(flat), flat.shape[0], _ptr(out)))
        return out.reshape(*lead, self.J)

    # -- fused (host numpy in/out) ------------------------------------------------------------------------
    def _frames(self, enc: np.ndarray):
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        if enc.ndim != 3:
            raise ValueError("enc must be [B,T,width]")
        B, T, W = enc.shape
        if W != self.J and W != self.E:
            raise ValueError(f"frame width {W} is neither joiner_dim {self.J} nor encoder_dim {self.E}")
        raw = 1 if (W == self.E and W != self.J) else 0   # E == J is ambiguous: pass enc_is_raw explicitly
        return enc, B, T, raw

    @staticmethod
    def _unpack(tokens, ts, n):
        return [tokens[b, :n[b]].tolist() for b in range(len(n))], [ts[b, :n[b]].tolist() for b in range(len(n))]

    def set_encoder_out_lens(self, lens: Optional[Sequence[int]]):
        """Per-stream frame counts (the seam's `encoder_out_lens`) for the NEXT greedy_offline / modified_beam_search call."""
        if lens is None:
            self._check(self._lib.k2b_set_encoder_out_lens(self._h, None, 0))
            return
        v = np.ascontiguousarray(lens, dtype=np.int64)
        self._check(self._lib.k2b_set_encoder_out_lens(self._h, _ptr(v), int(v.size)))

    def greedy_offline(self, enc: np.ndarray, mode: int, enc_is_raw: Optional[bool] = None, lens: Optional[Sequence[int]] = None):
        enc, B, T, raw = self._frames(enc)
        if enc_is_raw is not None:
            raw = int(enc_is_raw)
        if lens is not None:
            self.set_encoder_out_lens(lens)
        cap = max(T, 1)
        tokens = np.zeros((B, cap), np.int64); ts = np.zeros((B, cap), np.int32); n = np.zeros(B, np.int32)
        self._check(self._lib.k2b_greedy_offline(self._h, _ptr(enc), raw, B, T, mode, _ptr(tokens), _ptr(ts), _ptr(n), cap))
        return self._unpack(tokens, ts, n)

    def greedy_online_chunk(self, enc: np.ndarray, hyp: np.ndarray, enc_is_raw: Optional[bool] = None):
        enc, B, T, raw = self._frames(enc)
        if enc_is_raw is not None:
            raw = int(enc_is_raw)
        hyp = np.ascontiguousarray(hyp, dtype=np.int64).reshape(B, self.cfg.context_size).copy()
        cap = max(T, 1)
        tokens = np.zeros((B, cap), np.int64); ts = np.zeros((B, cap), np.int32); n = np.zeros(B, np.int32)
        self._check(self._lib.k2b_greedy_online_chunk(self._h, _ptr(enc), raw, B, T, _ptr(hyp), _ptr(tokens), _ptr(ts),
                                                      _ptr(n), cap))
        toks, tss = self._unpack(tokens, ts, n)
        return toks, tss, hyp

    def modified_beam_search(self, enc: np.ndarray, beam: int = 4, enc_is_raw: Optional[bool] = None,
                             lens: Optional[Sequence[int]] = None):
        enc, B, T, raw = self._frames(enc)
        if enc_is_raw is not None:
            raw = int(enc_is_raw)
        if lens is not None:
            self.set_encoder_out_lens(lens)
        cap = max(T, 1)
        tokens = np.zeros((B, cap), np.int64); ts = np.zeros((B, cap), np.int32); n = np.zeros(B, np.int32)
        score = np.zeros(B, np.float32)
        self._check(self._lib.k2b_modified_beam_search(self._h, _ptr(enc), raw, B, T, beam, _ptr(tokens), _ptr(ts),
                                                       _ptr(n), _ptr(score), cap))
        toks, tss = self._unpack(tokens, ts, n)
        return toks, tss, score

    # -- streaming modified_beam_search (hypotheses carried between chunks in device slots) ---------------------
    def beam_pool_create(self, max_streams: int, beam: int = 4, max_frames: int = 4096):
        self._check(self._lib.k2b_beam_pool_create(self._h, int(max_streams), int(beam), int(max_frames)))

    def beam_pool_reset(self, slot: int, hyp: Optional[Sequence[int]] = None):
        v = None if hyp is None else np.ascontiguousarray(hyp, dtype=np.int64)
        self._check(self._lib.k2b_beam_pool_reset(self._h, int(slot), _ptr(v)))

    def modified_beam_search_online_chunk(self, enc: np.ndarray, slots: Sequence[int], cap: int = 512,
                                          enc_is_raw: Optional[bool] = None):
        """Returns (tokens, timestamps, score, hyp): per stream the whole best hypothesis since the slot's reset."""
        enc, B, T, raw = self._frames(enc)
        if enc_is_raw is not None:
            raw = int(enc_is_raw)
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        assert sl.size == B
        tokens = np.zeros((B, cap), np.int64); ts = np.zeros((B, cap), np.int32); n = np.zeros(B, np.int32)
        score = np.zeros(B, np.float32); hyp = np.zeros((B, self.cfg.context_size), np.int64)
        self._check(self._lib.k2b_modified_beam_search_online_chunk(self._h, _ptr(enc), raw, B, T, _ptr(sl), _ptr(hyp), _ptr(tokens),
                                                                    _ptr(ts), _ptr(n), _ptr(score), cap))
        if (n > cap).any():
            raise K2bError(K2B_ERR_INVALID, f"cap {cap} is too small for a hypothesis of {int(n.max())} symbols")
        toks, tss = self._unpack(tokens, ts, n)
        return toks, tss, score, hyp

    def set_context_graph(self, graph=None):
        """graph: hotwords.ContextGraph (dense automaton over token ids) or None to clear."""
        if graph is None:
            self._check(self._lib.k2b_set_context_graph(self._h, None, None, None, 0))
            return
        nxt = np.ascontiguousarray(graph.next, dtype=np.int32)
        dlt = np.ascontiguousarray(graph.delta, dtype=np.float32)
        res = np.ascontiguousarray(graph.residual, dtype=np.float32)
        assert nxt.shape == dlt.shape == (res.size, self.V)
        self._check(self._lib.k2b_set_context_graph(self._h, _ptr(nxt), _ptr(dlt), _ptr(res), int(res.size)))

    def debug_backpointers(self, B: int, T: int, K: int) -> np.ndarray:
        """[B,T,K] back-pointer history of the last beam search: entry = (parent slot << 28) | (token + 1)."""
        out = np.zeros((B, T, K), np.int32)
        self._check(self._lib.k2b_debug_backpointers(self._h, _ptr(out), B, T, K))
        return out

    def ctc_greedy(self, logp: np.ndarray, blank: int = 0, frame_offset: Optional[Sequence[int]] = None,
                   prev: Optional[np.ndarray] = None, trailing_blank: Optional[np.ndarray] = None):
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        B, T, V = logp.shape
        cap = max(T, 1)
        tokens = np.zeros((B, cap), np.int64); ts = np.zeros((B, cap), np.int32); n = np.zeros(B, np.int32)
        fo = None if frame_offset is None else np.ascontiguousarray(frame_offset, dtype=np.int32)
        pv = None if prev is None else np.ascontiguousarray(prev, dtype=np.int64).copy()
        tb = None if trailing_blank is None else np.ascontiguousarray(trailing_blank, dtype=np.int32).copy()
        self._check(self._lib.k2b_ctc_greedy(self._h, _ptr(logp), B, T, V, blank, _ptr(fo), _ptr(pv), _ptr(tokens),
                                             _ptr(ts), _ptr(n), _ptr(tb), cap))
        toks, tss = self._unpack(tokens, ts, n)
        return toks, tss, tb, pv

    # -- on-device streaming state (stack_states / unstack_states) ------------------------------------------
    def state_pool_create(self, item_len: Sequence[int], max_streams: int):
        il = np.ascontiguousarray(item_len, dtype=np.int32)
        self._check(self._lib.k2b_state_pool_create(self._h, _ptr(il), int(il.size), int(max_streams)))
        self._pool_len = il.copy()

    def state_pool_stacked_floats(self, B: int) -> int:
        return int(self._lib.k2b_state_pool_stacked_floats(self._h, int(B)))

    def state_pool_put(self, slot: int, state: np.ndarray):
        st = np.ascontiguousarray(state, dtype=np.float32)
        assert st.size == int(self._pool_len.sum())
        self._check(self._lib.k2b_state_pool_put(self._h, int(slot), _ptr(st)))

    def state_pool_get(self, slot: int) -> np.ndarray:
        st = np.zeros(int(self._pool_len.sum()), np.float32)
        self._check(self._lib.k2b_state_pool_get(self._h, int(slot), _ptr(st)))
        return st

    def stack_states_dev(self, slots: Sequence[int], axis_len: Sequence[int], stacked_dev_ptr: int):
        sl = np.ascontiguousarray(slots, dtype=np.int32); ax = np.ascontiguousarray(axis_len, dtype=np.int32)
        self._check(self._lib.k2b_stack_states(self._h, _ptr(sl), int(sl.size), _ptr(ax), C.c_void_p(stacked_dev_ptr)))

    def unstack_states_dev(self, slots: Sequence[int], axis_len: Sequence[int], stacked_dev_ptr: int):
        sl = np.ascontiguousarray(slots, dtype=np.int32); ax = np.ascontiguousarray(axis_len, dtype=np.int32)
        self._check(self._lib.k2b_unstack_states(self._h, _ptr(sl), int(sl.size), _ptr(ax), C.c_void_p(stacked_dev_ptr)))

    # -- diagnostics ------------------------------------------------------------------------------------------
    def selftest_umma(self, A: np.ndarray, B: np.ndarray, mode: int, use_tma: bool = False) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
        N, K = B.shape
        D = np.zeros((128, N), np.float32)
        self._check(self._lib.k2b_selftest_umma(self._h, _ptr(A), _ptr(B), N, K, mode, int(use_tma), _ptr(D)))
        return D

    def debug_timeline(self):
        out = np.zeros((64, 148, 8), np.int64)
        self._check(self._lib.k2b_debug_timeline(self._h, _ptr(out)))
        return out

    def cluster_phase_cycles(self):
        out = np.zeros(20, np.int64)
        self._check(self._lib.k2b_cluster_phase_cycles(self._h, _ptr(out)))
        return out

    def selftest_umma2(self, A: np.ndarray, B: np.ndarray, ts: bool = False) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
        N, K = B.shape
        D = np.zeros((256, N), np.float32)
        self._check(self._lib.k2b_selftest_umma2(self._h, _ptr(A), _ptr(B), N, K, int(ts), _ptr(D)))
        return D

    def selftest_umma_bench(self, flavour: int, nkb: int = 4, reps: int = 8):
        out = np.zeros(2, np.int64)
        self._check(self._lib.k2b_selftest_umma_bench(self._h, flavour, nkb, reps, _ptr(out)))
        return out

    def selftest_collectives(self):
        out = np.zeros(6, np.int64)
        self._check(self._lib.k2b_selftest_collectives(self._h, _ptr(out)))
        return out

    def selftest_dsmem_bw(self, csize: int, nclusters: int, bytes_per_peer: int) -> int:
        out = np.zeros(1, np.int64)
        self._check(self._lib.k2b_selftest_dsmem_bw(self._h, csize, nclusters, bytes_per_peer, _ptr(out)))
        return int(out[0])

    def selftest_cluster(self, csize: int, nclusters: int):
        bad, done = C.c_int32(-1), C.c_int32(-1)
        self._check(self._lib.k2b_selftest_cluster(self._h, csize, nclusters, C.byref(bad), C.byref(done)))
        return int(bad.value), int(done.value)

    # -- page-locked host memory ---------------------------------------------------------------------------------
    @staticmethod
    def host_alloc(shape, dtype) -> np.ndarray:
        """A numpy array over page-locked memory from k2b_host_alloc (freed with host_free(arr))."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = _P()
        st = lib().k2b_host_alloc(C.byref(p), max(n, 1))
        if st != K2B_OK:
            raise K2bError(st, "k2b_host_alloc failed")
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        arr.flags.writeable = True
        _HOST_ALLOCS[arr.ctypes.data] = p.value
        return arr

    @staticmethod
    def host_free(arr: np.ndarray):
        p = _HOST_ALLOCS.pop(arr.ctypes.data, None)
        if p is not None:
            lib().k2b_host_free(p)

    # -- raw pointer access for device-resident / pinned buffers (bench.py, dist.py) ---------------------------
    def call(self, name: str, *args):
        """Call an entry point with raw addresses (ints / torch tensors / numpy arrays / None)."""
        fn = getattr(self._lib, name)
        conv = [(_ptr(a) if (a is None or isinstance(a, np.ndarray) or hasattr(a, "data_ptr")) else a) for a in args]
        self._check(fn(self._h, *conv))
