"""Host-side mirror of the reference's proj seam (IOfflineProj / IOnlineProj) over libk2b200.so.

The reference host language is C# and no .NET toolchain exists in this image, so the host side above the
C ABI is written in Python with the reference's own member names, argument meaning and error behaviour;
the C# classes a maintainer would add (`OfflineProjOfB200 : IOfflineProj`, `OnlineProjOfB200 : IOnlineProj`)
are listed in INTEGRATION.md and map 1:1 onto these.

Scope note: the encoder network is out of scope; `EncoderProj` therefore receives
the encoder's output frames in `Speech` instead of fbank features and only applies the encoder_proj
Linear (or nothing when the frames are already projected / are CTC log-probs).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _native
from .synth import ModelDims


# ---- payload types of the seam (ref Model/*.cs) ---------------------------------------------------------
@dataclass
class OfflineCustomMetadata:            # ref Model/OfflineCustomMetadata.cs:16-33
    Version: Optional[str] = None
    Model_type: Optional[str] = "zipformer2"
    Model_author: Optional[str] = None
    Context_size: int = 2
    Vocab_size: int = 500
    Joiner_dim: int = 512
    Comment: Optional[str] = None
    Feature_type: str = "fbank"


@dataclass
class OnlineCustomMetadata(OfflineCustomMetadata):   # ref Model/OnlineCustomMetadata.cs (search-relevant fields)
    T: int = 8                          # encoder frames per chunk on our path
    Decode_chunk_len: int = 8


@dataclass
class OfflineInputEntity:               # ref Model/OfflineInputEntity.cs
    Speech: Optional[np.ndarray] = None
    SpeechLength: int = 0


@dataclass
class OnlineInputEntity:                # ref Model/OnlineInputEntity.cs
    Speech: Optional[np.ndarray] = None
    SpeechLength: int = 0


@dataclass
class EncoderOutputEntity:              # ref Model/EncoderOutputEntity.cs:10-20
    encoder_out: Optional[np.ndarray] = None
    encoder_out_lens: Optional[np.ndarray] = None
    encoder_out_states: Optional[list] = None


@dataclass
class DecoderOutputEntity:              # ref Model/DecoderOutputEntity.cs:6-10
    decoder_out: Optional[np.ndarray] = None


@dataclass
class JoinerOutputEntity:               # ref Model/JoinerOutputEntity.cs:8-16
    Logit: Optional[np.ndarray] = None   # flat copy
    Logits: Optional[np.ndarray] = None  # [N,V] view


def _pad_batch(inputs, width: int) -> np.ndarray:
    """Right-pad every stream's frames with zeros to the longest (the reference pads its encoder input the
    same way, ref Utils/PadHelper.cs, and then decodes the padding: Q7)."""
    rows = [np.asarray(x.Speech, np.float32).reshape(-1, width) for x in inputs]
    T = max((r.shape[0] for r in rows), default=0)
    out = np.zeros((len(rows), T, width), np.float32)
    for i, r in enumerate(rows):
        out[i, :r.shape[0]] = r
    return out


class _ProjBase:
    """What OfflineProjOfTransducer and the five OnlineProjOf* classes share on this path."""

    def __init__(self, dims: ModelDims, weights: Optional[dict], model_type: str, device: int = 0,
                 precision: str = "fp32", neg_id_wrap: bool = False, frames_are_raw: Optional[bool] = None):
        self._dims = dims
        self._native = _native.Handle(
            vocab_size=dims.vocab_size, joiner_dim=dims.joiner_dim, decoder_dim=dims.decoder_dim,
            encoder_dim=dims.encoder_dim, context_size=dims.context_size, blank_id=dims.blank_id,
            sos_eos_id=dims.sos_eos_id, unk_id=dims.unk_id, device=device,
            neg_id_mode=_native.NEGID_WRAP if neg_id_wrap else _native.NEGID_MASK,
            precision=_native.PREC_NAMES[precision])
        if weights is not None:
            self._native.load_weights(weights)
        self._frames_are_raw = (dims.encoder_dim > 0) if frames_are_raw is None else frames_are_raw
        # the interface exposes ORT sessions (ref IOfflineProj.cs:8-22); there are none on this path
        self.EncoderSession = None
        self.DecoderSession = None
        self.JoinerSession = None
        self.Blank_id = dims.blank_id          # ref OfflineModel.cs:18-20
        self.Sos_eos_id = dims.sos_eos_id
        self.Unk_id = dims.unk_id
        self._model_type = model_type
        self._disposed = False

    @property
    def Native(self) -> _native.Handle:
        return self._native

    @property
    def FrameWidth(self) -> int:
        return self._dims.encoder_dim if self._frames_are_raw else self._dims.joiner_dim

    def DecoderProj(self, decoder_input: Optional[np.ndarray], batchSize: int) -> DecoderOutputEntity:
        """ref OfflineProjOfTransducer.cs:93-123 (null input -> batchSize x {-1, blank}, :97-110)."""
        out = self._native.decoder_proj(decoder_input, batchSize)
        return DecoderOutputEntity(decoder_out=out.reshape(-1))

    def JoinerProj(self, encoder_out: np.ndarray, decoder_out: np.ndarray) -> JoinerOutputEntity:
        """ref OfflineProjOfTransducer.cs:125-152."""
        logits = self._native.joiner_proj(encoder_out, decoder_out)
        return JoinerOutputEntity(Logit=logits.reshape(-1), Logits=logits)

    def Dispose(self):                        # ref OfflineProjOfTransducer.cs:154-189
        if not self._disposed:
            self._native.close()
            self._disposed = True


class OfflineProjOfB200(_ProjBase):
    """IOfflineProj (ref IOfflineProj.cs:6-48) over libk2b200.so: takes the place of OfflineProjOfTransducer."""

    def __init__(self, dims: ModelDims, weights: dict, model_type: str = "zipformer2", **kw):
        super().__init__(dims, weights, model_type, **kw)
        self.CustomMetadata = OfflineCustomMetadata(Model_type=model_type, Context_size=dims.context_size,
                                                    Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim)

    def EncoderProj(self, modelInputs: List[OfflineInputEntity], batchSize: int) -> EncoderOutputEntity:
        """ref OfflineProjOfTransducer.cs:48-92: returns flat [B,T,J] frames + lens (never read, Q7)."""
        try:
            x = _pad_batch(modelInputs, self.FrameWidth)
            if self._frames_are_raw:
                x = self._native.encoder_proj(x)
            lens = np.full(batchSize, x.shape[1], np.int64)
            return EncoderOutputEntity(encoder_out=x.reshape(-1), encoder_out_lens=lens)
        except Exception as ex:
            raise Exception("EncoderProj failed") from ex   # ref :87-90


class OfflineProjOfB200ctc(_ProjBase):
    """CTC variant of the seam (ref OfflineProjOfZipformer2ctc.cs): EncoderProj returns [B,T,V] log-probs,
    DecoderProj / JoinerProj return null (ref :93-101). Needs no weights."""

    def __init__(self, dims: ModelDims, **kw):
        super().__init__(dims, None, "zipformer2ctc", frames_are_raw=False, **kw)
        self.CustomMetadata = OfflineCustomMetadata(Model_type="zipformer2ctc", Context_size=dims.context_size,
                                                    Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim)

    @property
    def FrameWidth(self) -> int:
        return self._dims.vocab_size

    def EncoderProj(self, modelInputs, batchSize: int, statesList=None) -> EncoderOutputEntity:
        x = _pad_batch(modelInputs, self._dims.vocab_size)
        return EncoderOutputEntity(encoder_out=x.reshape(-1), encoder_out_lens=np.full(batchSize, x.shape[1], np.int64),
                                   encoder_out_states=statesList)

    def DecoderProj(self, decoder_input, batchSize):
        return None

    def JoinerProj(self, encoder_out, decoder_out):
        return None


@dataclass
class DeviceStates:
    """One stream's encoder caches, resident in a slot of the library's device pool (k2b_state_pool_*): what takes the place of
    the List<List<float[]>> of ref OnlineStream.cs:14 when the caches never leave the GPU."""
    slot: int


@dataclass
class StackedStates:
    """The batched caches of one encoder call (ref stack_states' return value), as ONE device buffer laid out as k2b_stack_states
    documents (tensor i of all B streams at float offset B * off_i), plus the slots they came from."""
    buffer: object            # torch tensor on the handle's device
    slots: List[int]


def zipformer2_state_layout(num_encoder_layers, encoder_dims, num_heads, query_head_dims, value_head_dims, left_context_len,
                            cnn_module_kernels):
    """Per-stream cache tensors of a streaming zipformer2 from its ONNX metadata (keys of ref Model/OnlineCustomMetadata.cs), in the
    order of ref OnlineProjOfZipformer2.cs:63-111 (per layer: key, nonlin_attn, val1, val2, conv1, conv2; then embed_states and
    processed_lens), with the "axisnum" of stack_states for each (ref :236-341). Returns (item_len, axis_len)."""
    il, ax = [], []
    for i, nl in enumerate(num_encoder_layers):
        d = encoder_dims[i]
        key, val = query_head_dims[i] * num_heads[i], value_head_dims[i] * num_heads[i]
        left, pad = left_context_len[i], cnn_module_kernels[i] // 2
        for _ in range(nl):
            il += [left * key, left * (3 * d // 4), left * val, left * val, d * pad, d * pad]
            ax += [key, left, val, val, d * pad, d * pad]
    il += [128 * 3 * 19, 1]           # embed_states (ref :58-61), processed_lens
    ax += [128 * 3 * 19, 1]
    return il, ax


class OnlineProjOfB200(_ProjBase):
    """IOnlineProj (ref IOnlineProj.cs:8-72) over libk2b200.so: stands in for OnlineProjOfZipformer /
    Zipformer2 / Lstm / Conformer, whose DecoderProj/JoinerProj bodies are identical.

    Encoder caches: with `state_layout` = (item_len, axis_len) the caches live in the library's device pool and
    GetEncoderInitStates / stack_states / unstack_states are k2b_state_pool_put / k2b_stack_states / k2b_unstack_states (one launch
    each, nothing crosses PCIe) - the O(B * cache) Array.Copy loops of ref OnlineProjOfZipformer2.cs:144-489 on the device. Without a
    layout the caches are opaque host objects passed through (the encoder network itself is out of scope either way; `encoder_hook`
    is where it would run: it receives the stacked device buffer and updates it in place)."""

    def __init__(self, dims: ModelDims, weights: dict, model_type: str = "zipformer2", chunk_frames: int = 8,
                 state_layout=None, max_streams: int = 512, encoder_hook=None, **kw):
        super().__init__(dims, weights, model_type, **kw)
        self.CustomMetadata = OnlineCustomMetadata(Model_type=model_type, Context_size=dims.context_size,
                                                   Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim,
                                                   T=chunk_frames, Decode_chunk_len=chunk_frames)
        self.ChunkLength = chunk_frames        # ref OnlineModel.cs:48 (here counted in encoder frames)
        self.ShiftLength = chunk_frames        # ref OnlineModel.cs:49
        self.FeatureDim = self.FrameWidth
        self.SampleRate = 16000
        self._layout = None
        self._encoder_hook = encoder_hook
        self._device = kw.get("device", 0)
        if state_layout is not None:
            item_len, axis_len = state_layout
            self._layout = (np.asarray(item_len, np.int32), np.asarray(axis_len, np.int32))
            self._native.state_pool_create(item_len, max_streams)
            self._free_state_slots = list(range(max_streams - 1, -1, -1))
            self._zeros = np.zeros(int(self._layout[0].sum()), np.float32)

    def GetEncoderInitStates(self, batchSize: int = 1):
        """ref OnlineProjOfZipformer2.cs:63-111: all-zero caches of one stream - here written into a fresh pool slot."""
        if self._layout is None:
            return [[] for _ in range(batchSize)]
        if not self._free_state_slots:
            raise Exception("OnlineProjOfB200: more concurrent streams than max_streams")
        slot = self._free_state_slots.pop()
        self._native.state_pool_put(slot, self._zeros)
        return DeviceStates(slot)

    def ReleaseStates(self, states):
        if self._layout is not None and isinstance(states, DeviceStates):
            self._free_state_slots.append(states.slot)

    def stack_states(self, stateList):
        """ref OnlineProjOfZipformer2.cs:144-362: per-stream caches -> batched tensors (batch on axis 1)."""
        if self._layout is None:
            return stateList
        import torch
        slots = [st.slot for st in stateList]
        n = self._native.state_pool_stacked_floats(len(slots))
        buf = torch.empty(n, dtype=torch.float32, device=f"cuda:{self._device}")
        torch.cuda.current_stream(self._device).synchronize()       # the buffer is written on the handle's stream
        self._native.stack_states_dev(slots, self._layout[1], buf.data_ptr())
        return StackedStates(buf, slots)

    def unstack_states(self, encoder_out_states):
        """ref OnlineProjOfZipformer2.cs:363-489: the encoder's new batched caches -> back into every stream's slot."""
        if self._layout is None:
            return encoder_out_states
        st = encoder_out_states
        self._native.unstack_states_dev(st.slots, self._layout[1], st.buffer.data_ptr())
        self._native.sync()
        return [DeviceStates(s) for s in st.slots]

    def EncoderProj(self, modelInputs: List[OnlineInputEntity], batchSize: int, statesList=None) -> EncoderOutputEntity:
        """ref OnlineProjOfZipformer2.cs:491-618 reduced to its tail: [B,T',J] projected frames + caches."""
        try:
            x = _pad_batch(modelInputs, self.FrameWidth)
            if self._frames_are_raw:
                x = self._native.encoder_proj(x)
            if self._encoder_hook is not None and isinstance(statesList, StackedStates):
                self._native.sync()
                self._encoder_hook(statesList.buffer)            # the encoder network would consume / produce the caches here
            return EncoderOutputEntity(encoder_out=x.reshape(-1), encoder_out_lens=np.full(batchSize, x.shape[1], np.int64),
                                       encoder_out_states=statesList)
        except Exception as ex:
            raise Exception("EncoderProj failed") from ex


class OnlineProjOfB200ctc(OfflineProjOfB200ctc):
    """ref OnlineProjOfZipformer2ctc.cs: as the offline CTC proj plus the chunk geometry."""

    def __init__(self, dims: ModelDims, chunk_frames: int = 8, **kw):
        super().__init__(dims, **kw)
        self.CustomMetadata = OnlineCustomMetadata(Model_type="zipformer2ctc", Context_size=dims.context_size,
                                                   Vocab_size=dims.vocab_size, Joiner_dim=dims.joiner_dim,
                                                   T=chunk_frames, Decode_chunk_len=chunk_frames)
        self.ChunkLength = chunk_frames
        self.ShiftLength = chunk_frames
        self.FeatureDim = dims.vocab_size
        self.SampleRate = 16000

    def GetEncoderInitStates(self, batchSize: int = 1):
        return [[] for _ in range(batchSize)]

    def stack_states(self, stateList):
        return stateList

    def unstack_states(self, encoder_out_states):
        return encoder_out_states

    def ReleaseStates(self, states):
        pass
