"""Host-side mirror of OfflineRecognizer / OnlineRecognizer / OfflineStream / OnlineStream.

Same public names and result conventions as the reference (ref OfflineRecognizer.cs, OnlineRecognizer.cs,
OfflineStream.cs, OnlineStream.cs). Two implementations of every search delegate:

  fused=True  (default)  one call into libk2b200.so per GetResult(s): k2b_greedy_offline,
                         k2b_greedy_online_chunk, k2b_modified_beam_search, k2b_ctc_greedy;
  fused=False            the reference's own per-frame loop shape, with DecoderProj / JoinerProj served by
                         the fine-grained entry points (what a drop-in that only swaps the proj gives).

Both leave exactly what the reference leaves in stream.Tokens / Timestamps / Hyp / NumTrailingBlank,
including the list seeds ({-1, blank}, the 2B blanks of Q5, ...). The feature frontend and the encoder
network are out of scope: streams accept encoder output frames (AcceptFrames) where the reference
accepts audio samples (AddSamples).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native
from .text import decode_multi
from .proj import (OfflineInputEntity, OnlineInputEntity, OfflineProjOfB200, OfflineProjOfB200ctc,
                   OnlineProjOfB200, OnlineProjOfB200ctc)


@dataclass
class OfflineRecognizerResultEntity:     # ref Model/OfflineRecognizerResultEntity.cs
    text: str = ""
    tokens: List[str] = field(default_factory=list)
    timestamps: List[int] = field(default_factory=list)


OnlineRecognizerResultEntity = OfflineRecognizerResultEntity


def _argmax_hi_rows(logits: np.ndarray) -> np.ndarray:
    """Host argmax of the fine-grained path: ties -> larger index (ref OfflineRecognizer.cs:150-154)."""
    V = logits.shape[1]
    return V - 1 - np.argmax(logits[:, ::-1], axis=1)


def _decode_text(token_ids: Sequence[int], symbols: Optional[Sequence[str]], online: bool = False):
    """DecodeMulti of one stream (ref OfflineRecognizer.cs:432-467, OnlineRecognizer.cs:321-351): text.py restates it in full -
    stop at id 2, <0xNN> byte fallback, byte-level BPE, lower-casing."""
    return decode_multi(token_ids, symbols, online=online)


# ---- streams ------------------------------------------------------------------------------------------
class OfflineStream:
    """ref OfflineStream.cs:7-68. Tokens start {blank, blank} (ref :34)."""

    def __init__(self, context_size: int = 2, blank_id: int = 0):
        self._context_size = context_size
        self.OfflineInputEntity = OfflineInputEntity()
        self.Tokens: List[int] = [blank_id, blank_id]
        self.Timestamps: List[int] = []
        self.FrameOffset = 0
        self.NumTrailingBlank = 0

    def AcceptFrames(self, frames: np.ndarray):
        """Stands where AddSamples stands (ref :43-57): appends encoder output frames [t,width]."""
        f = np.asarray(frames, np.float32)
        f = f.reshape(-1, f.shape[-1])
        if self.OfflineInputEntity.Speech is None or self.OfflineInputEntity.SpeechLength == 0:
            self.OfflineInputEntity.Speech = f.copy()
        else:
            self.OfflineInputEntity.Speech = np.concatenate([self.OfflineInputEntity.Speech, f], axis=0)
        self.OfflineInputEntity.SpeechLength = self.OfflineInputEntity.Speech.size

    def RemoveSamples(self):                 # ref :58-68
        if len(self.Tokens) > self._context_size:
            self.OfflineInputEntity.Speech = None
            self.OfflineInputEntity.SpeechLength = 0


class OnlineStream:
    """ref OnlineStream.cs:7-161. Hyp and Tokens start {blank, blank} (ref :44-45)."""

    def __init__(self, onlineProj):
        self._chunkLength = onlineProj.ChunkLength
        self._shiftLength = onlineProj.ShiftLength
        self._width = onlineProj.FrameWidth
        blank = onlineProj.Blank_id
        self.OnlineInputEntity = OnlineInputEntity()
        self.Hyp = np.array([blank, blank], np.int64)
        self.Tokens: List[int] = [blank, blank]
        self.Timestamps: List[int] = []
        self.States = onlineProj.GetEncoderInitStates()
        self.FrameOffset = 0
        self.NumTrailingBlank = 0
        self._finished = False
        self.BeamSlot: Optional[int] = None      # streaming modified_beam_search: this stream's slot in the device beam pool
        self.NumProcessedFrames = 0              # encoder frames decoded since the stream began (or was last reset)

    def AcceptFrames(self, frames: np.ndarray):
        f = np.asarray(frames, np.float32).reshape(-1, self._width)
        cur = self.OnlineInputEntity.Speech
        self.OnlineInputEntity.Speech = f.copy() if cur is None or len(cur) == 0 else np.concatenate([cur, f], axis=0)
        self.OnlineInputEntity.SpeechLength = self.OnlineInputEntity.Speech.size

    def InputFinished(self):
        self._finished = True

    def GetDecodeChunk(self) -> Optional[np.ndarray]:
        """ref :82-101: the next ChunkLength frames, or null when fewer are buffered."""
        sp = self.OnlineInputEntity.Speech
        if sp is None or sp.shape[0] < self._chunkLength:
            return None
        return sp[: self._chunkLength]

    def RemoveChunk(self):                   # ref :103-117
        sp = self.OnlineInputEntity.Speech
        if sp is not None and sp.shape[0] >= self._shiftLength:
            self.OnlineInputEntity.Speech = sp[self._shiftLength:]
            self.OnlineInputEntity.SpeechLength = self.OnlineInputEntity.Speech.size

    def IsFinished(self) -> bool:
        sp = self.OnlineInputEntity.Speech
        return self._finished and (sp is None or sp.shape[0] < self._chunkLength)


# ---- offline ---------------------------------------------------------------------------------------------
class OfflineRecognizer:
    """ref OfflineRecognizer.cs:12-91. `decodingMethod`: greedy_search | greedy_search_ctc |
    modified_beam_search (the last does not exist in the reference, whose switch falls back to greedy:
    ref :54-68). batch_mode: 'compat' restates the reference batch loop incl. Q5/Q6, 'per_stream' is ours."""

    def __init__(self, offlineProj, tokens: Optional[Sequence[str]] = None, decodingMethod: str = "greedy_search",
                 fused: bool = True, maxActivePaths: int = 4, batch_mode: str = "compat", maxSymPerFrame: int = 1):
        self._offlineProj = offlineProj
        self._max_sym_per_frame = maxSymPerFrame                         # ref :19 (fixed to 1 there)
        if maxSymPerFrame != 1:
            offlineProj.Native.set_option("max_sym_per_frame", maxSymPerFrame)
        self._tokens = list(tokens) if tokens is not None else None
        self._blank_id = offlineProj.Blank_id
        self._unk_id = offlineProj.Unk_id
        self._fused = fused
        self._beam = maxActivePaths
        self._batch_mode = batch_mode
        if offlineProj.CustomMetadata.Model_type == "zipformer2ctc":     # ref :46-49
            decodingMethod = "greedy_search_ctc"
        if decodingMethod == "greedy_search_ctc":
            self._forward, self._forwardBatch = self.ForwardGreedySearchCTC, self.ForwardBatchGreedySearchCTC
        elif decodingMethod == "modified_beam_search":
            self._forward, self._forwardBatch = self.ForwardModifiedBeamSearch, self.ForwardBatchModifiedBeamSearch
        else:                                                            # ref :64-67 default -> greedy
            self._forward, self._forwardBatch = self.ForwardGreedySearch, self.ForwardBatchGreedySearch
        self.LastScores: List[float] = []

    def CreateOfflineStream(self) -> OfflineStream:                      # ref :71-75
        return OfflineStream(self._offlineProj.CustomMetadata.Context_size, self._blank_id)

    def GetResult(self, stream: OfflineStream) -> OfflineRecognizerResultEntity:   # ref :77-83
        self._forward(stream)
        return self.DecodeMulti([stream])[0]

    def GetResults(self, streams: List[OfflineStream]) -> List[OfflineRecognizerResultEntity]:   # ref :85-91
        self._forwardBatch(streams)
        return self.DecodeMulti(streams)

    def DecodeMulti(self, streams):
        out = []
        for s in streams:
            text, toks = _decode_text(s.Tokens, self._tokens)
            out.append(OfflineRecognizerResultEntity(text=text, tokens=toks, timestamps=list(s.Timestamps)))
        return out

    # -- greedy, single stream (ref :93-187) -------------------------------------------------------------------
    def ForwardGreedySearch(self, stream: OfflineStream):
        try:
            proj = self._offlineProj
            enc = proj.EncoderProj([stream.OfflineInputEntity], 1)
            J = proj.CustomMetadata.Joiner_dim
            frames = enc.encoder_out.reshape(1, -1, J)
            if self._fused:
                toks, tss = proj.Native.greedy_offline(frames, _native.GREEDY_SINGLE, enc_is_raw=False)
                stream.Tokens = [-1, self._blank_id] + toks[0]
                stream.Timestamps.extend(tss[0])
                return
            TT = frames.shape[1]
            ctx = proj.CustomMetadata.Context_size
            hypList = [-1, self._blank_id]
            decoder_out = proj.DecoderProj(np.array(hypList, np.int64), 1).decoder_out
            timestamp: List[int] = []
            t, sym_per_utt, sym_per_frame = 0, 0, 0
            while t < TT and sym_per_utt < 1000:
                if sym_per_frame >= self._max_sym_per_frame:             # ref :129-134
                    sym_per_frame = 0
                    t += 1
                    continue
                logits = proj.JoinerProj(frames[0, t], decoder_out).Logits
                y = int(_argmax_hi_rows(logits)[0])
                if y != self._blank_id and y != self._unk_id:
                    hypList.append(y)
                    timestamp.append(t)
                    decoder_out = proj.DecoderProj(np.array(hypList[-ctx:], np.int64), 1).decoder_out
                    sym_per_utt += 1
                    sym_per_frame += 1
                else:                                                    # ref :174-178
                    sym_per_frame = 0
                    t += 1
            stream.Tokens = hypList
            stream.Timestamps.extend(timestamp)
        except Exception as ex:
            raise Exception("Offline recognition failed") from ex       # ref :183-186

    # -- greedy, batch (ref :189-303) --------------------------------------------------------------------------
    def ForwardBatchGreedySearch(self, streams: List[OfflineStream]):
        try:
            proj = self._offlineProj
            B = len(streams)
            if B == 0:
                return
            enc = proj.EncoderProj([s.OfflineInputEntity for s in streams], B)
            J = proj.CustomMetadata.Joiner_dim
            frames = enc.encoder_out.reshape(B, -1, J)
            T = frames.shape[1]
            compat = self._batch_mode == "compat"
            seed_tok = [self._blank_id] * (2 * B) if compat else [-1, self._blank_id]     # Q5 (ref :250-267)
            seed_ts = [0] * (2 * B) if compat else []
            if self._fused:
                mode = _native.GREEDY_BATCH_COMPAT if compat else _native.GREEDY_PER_STREAM
                toks, tss = proj.Native.greedy_offline(frames, mode, enc_is_raw=False)
            else:
                ctx = proj.CustomMetadata.Context_size
                decoder_out = proj.DecoderProj(None, B).decoder_out
                tokens = [list(seed_tok) for _ in range(B)]
                tsl = [[] for _ in range(B)]
                for t in range(T):
                    logits = proj.JoinerProj(frames[:, t, :], decoder_out).Logits
                    ys = _argmax_hi_rows(logits)
                    emitted = False
                    for m in range(B):
                        if ys[m] != self._blank_id and ys[m] != self._unk_id:
                            tokens[m].append(int(ys[m]))
                            tsl[m].append(t)
                            emitted = True
                    if emitted:
                        if compat:
                            di = np.array([tk[-ctx:] for tk in tokens], np.int64)
                        else:
                            di = np.array([tk[-ctx:] for tk in tokens], np.int64)
                        decoder_out = proj.DecoderProj(di, B).decoder_out
                toks = [tk[len(seed_tok):] for tk in tokens]
                tss = tsl
            for m, s in enumerate(streams):
                if T == 0 and compat:
                    continue                 # Q13: the reference leaves null here; we leave the stream untouched
                s.Tokens = list(seed_tok) + toks[m]
                s.Timestamps.extend(seed_ts + tss[m])
                s.RemoveSamples()
        except Exception as ex:
            raise Exception("Offline recognition failed") from ex       # ref :299-302

    # -- CTC (ref :305-430) ------------------------------------------------------------------------------------
    def ForwardGreedySearchCTC(self, stream: OfflineStream):
        try:
            proj = self._offlineProj
            enc = proj.EncoderProj([stream.OfflineInputEntity], 1)
            V = len(self._tokens) if self._tokens is not None else proj.CustomMetadata.Vocab_size   # ref :325
            logp = enc.encoder_out.reshape(1, -1, V)
            toks, tss, tb, _ = proj.Native.ctc_greedy(logp, blank=proj.Blank_id, trailing_blank=np.zeros(1, np.int32))
            stream.NumTrailingBlank = int(tb[0])
            stream.Tokens = [-1, self._blank_id] + toks[0]               # ref :318-320
            stream.Timestamps.extend(tss[0])
        except Exception:
            pass                                                         # Q11: ref :359-362 swallows

    def ForwardBatchGreedySearchCTC(self, streams: List[OfflineStream]):
        try:
            proj = self._offlineProj
            B = len(streams)
            if B == 0:
                return
            enc = proj.EncoderProj([s.OfflineInputEntity for s in streams], B)
            V = len(self._tokens) if self._tokens is not None else proj.CustomMetadata.Vocab_size
            logp = enc.encoder_out.reshape(B, -1, V)
            fo = np.array([s.FrameOffset for s in streams], np.int32)
            tb = np.array([s.NumTrailingBlank for s in streams], np.int32)
            toks, tss, tb, _ = proj.Native.ctc_greedy(logp, blank=proj.Blank_id, frame_offset=fo, trailing_blank=tb)
            for m, s in enumerate(streams):
                s.Tokens = list(s.Tokens) + toks[m]                      # ref :381, :409
                s.Timestamps = list(s.Timestamps) + tss[m]
                s.NumTrailingBlank = int(tb[m])
                s.RemoveSamples()
        except Exception as ex:
            raise Exception("Speech recognition failed") from ex        # ref :426-429

    # -- modified_beam_search (ours; icefall semantics) -------------------------------------------------------
    def ForwardModifiedBeamSearch(self, stream: OfflineStream):
        self.ForwardBatchModifiedBeamSearch([stream])

    def ForwardBatchModifiedBeamSearch(self, streams: List[OfflineStream]):
        try:
            proj = self._offlineProj
            B = len(streams)
            if B == 0:
                return
            enc = proj.EncoderProj([s.OfflineInputEntity for s in streams], B)
            J = proj.CustomMetadata.Joiner_dim
            frames = enc.encoder_out.reshape(B, -1, J)
            toks, tss, score = proj.Native.modified_beam_search(frames, self._beam, enc_is_raw=False)
            self.LastScores = [float(x) for x in score]
            for m, s in enumerate(streams):
                s.Tokens = [-1, self._blank_id] + toks[m]
                s.Timestamps.extend(tss[m])
                s.RemoveSamples()
        except Exception as ex:
            raise Exception("Offline recognition failed") from ex

    def Dispose(self):                                                   # ref :579-600
        self._offlineProj.Dispose()


# ---- online ----------------------------------------------------------------------------------------------
class OnlineRecognizer:
    """ref OnlineRecognizer.cs:11-84. `decodingMethod`: greedy_search | greedy_search_ctc | modified_beam_search - the reference
    accepts the last together with `maxActivePaths` and silently decodes greedily (ref :18-19, :46-57); here it selects the
    streaming beam search of the library (hypotheses carried between chunks in device slots). `enableEndpoint` (accepted and
    ignored by the reference, ref :19) switches the NumTrailingBlank-based endpoint rules on (IsEndpoint)."""

    # endpoint rules [EXT: sherpa-onnx defaults], in encoder frames of 40 ms: trailing silence with nothing decoded, trailing
    # silence after something was decoded, utterance length
    RULE1_FRAMES, RULE2_FRAMES, RULE3_FRAMES = 60, 30, 500

    def __init__(self, onlineProj, tokens: Optional[Sequence[str]] = None, decodingMethod: str = "greedy_search",
                 fused: bool = True, maxActivePaths: int = 4, enableEndpoint: int = 0, maxStreams: int = 512,
                 maxFrames: int = 4096):
        self._onlineProj = onlineProj
        self._tokens = list(tokens) if tokens is not None else None
        self._fused = fused
        self._beam = maxActivePaths
        self._enableEndpoint = enableEndpoint
        self._free_slots: List[int] = []
        if onlineProj.CustomMetadata.Model_type == "zipformer2ctc":      # ref :34-37
            decodingMethod = "greedy_search_ctc"
        if decodingMethod == "greedy_search_ctc":
            self._forwardBatch = self.ForwardBatchGreedySearchCTC
        elif decodingMethod == "modified_beam_search":
            self._forwardBatch = self.ForwardBatchModifiedBeamSearch
            onlineProj.Native.beam_pool_create(maxStreams, maxActivePaths, maxFrames)
            self._free_slots = list(range(maxStreams - 1, -1, -1))
            self._cap = maxFrames
        else:
            self._forwardBatch = self.ForwardBatchGreedySearch           # ref :46-57
        self._decodingMethod = decodingMethod

    def CreateOnlineStream(self) -> OnlineStream:                        # ref :60-64
        s = OnlineStream(self._onlineProj)
        if self._decodingMethod == "modified_beam_search":
            if not self._free_slots:
                raise Exception("OnlineRecognizer: more concurrent streams than maxStreams")
            s.BeamSlot = self._free_slots.pop()
            self._onlineProj.Native.beam_pool_reset(s.BeamSlot, s.Hyp)
        return s

    def ReleaseOnlineStream(self, stream: OnlineStream):
        """Gives the stream's device slots (beam pool, encoder caches) back; the reference leaves this to the GC."""
        if stream.BeamSlot is not None:
            self._free_slots.append(stream.BeamSlot)
            stream.BeamSlot = None
        self._onlineProj.ReleaseStates(stream.States)

    def GetResult(self, stream: OnlineStream):                           # ref :66-74
        return self.GetResults([stream])[0]

    def GetResults(self, streams: List[OnlineStream]):                   # ref :76-84
        self._forwardBatch(streams)
        out = []
        for s in streams:
            text, toks = _decode_text(s.Tokens, self._tokens, online=True)
            out.append(OnlineRecognizerResultEntity(text=text, tokens=toks, timestamps=list(s.Timestamps)))
        return out

    # -- endpointing on NumTrailingBlank (the field exists in the reference, OnlineStream.cs:16,55,
    #    but nothing produces or consumes it on the transducer path) --------------------------------------------------------------
    def IsEndpoint(self, stream: OnlineStream) -> bool:
        if not self._enableEndpoint:
            return False
        decoded = len(stream.Tokens) > self._onlineProj.CustomMetadata.Context_size
        if not decoded and stream.NumTrailingBlank >= self.RULE1_FRAMES:
            return True
        if decoded and stream.NumTrailingBlank >= self.RULE2_FRAMES:
            return True
        return stream.NumProcessedFrames >= self.RULE3_FRAMES

    def Reset(self, stream: OnlineStream):
        """After an endpoint: a new utterance on the same stream (encoder caches are kept, the search state starts over)."""
        blank = self._onlineProj.Blank_id
        stream.Hyp = np.array([blank, blank], np.int64)
        stream.Tokens = [blank, blank]
        stream.Timestamps = []
        stream.NumTrailingBlank = 0
        stream.NumProcessedFrames = 0
        if stream.BeamSlot is not None:
            self._onlineProj.Native.beam_pool_reset(stream.BeamSlot, stream.Hyp)

    @staticmethod
    def _count_trailing(stream: OnlineStream, n_frames: int, chunk_ts: Sequence[int]):
        """NumTrailingBlank from one chunk's chunk-local timestamps (the rule of the CTC loop, ref OfflineRecognizer.cs:337-344)."""
        stream.NumProcessedFrames += n_frames
        if len(chunk_ts):
            stream.NumTrailingBlank = n_frames - 1 - int(chunk_ts[-1])
        else:
            stream.NumTrailingBlank += n_frames

    def _collect(self, streams: List[OnlineStream]):
        """ref :97-120: streams without a full chunk are REMOVED from the caller's list."""
        inputs, active, skipped = [], [], []
        for s in streams:
            chunk = s.GetDecodeChunk()
            if chunk is None:
                skipped.append(s)
                continue
            inputs.append(OnlineInputEntity(Speech=chunk.copy(), SpeechLength=chunk.size))
            s.RemoveChunk()
            active.append(s)
        if not inputs:
            return [], []
        for s in skipped:
            streams.remove(s)
        return inputs, active

    def ForwardBatchGreedySearch(self, streams: List[OnlineStream]):     # ref :85-219
        if len(streams) == 0:
            return
        inputs, active = self._collect(streams)
        if not inputs:
            return
        proj = self._onlineProj
        B = len(inputs)
        ctx = proj.CustomMetadata.Context_size
        try:
            states = proj.stack_states([s.States for s in active])
            enc = proj.EncoderProj(inputs, B, states)
            J = proj.CustomMetadata.Joiner_dim
            frames = enc.encoder_out.reshape(B, -1, J)
            hyps = np.stack([np.asarray(s.Hyp, np.int64) for s in active])
            if self._fused:
                toks, tss, hyp_out = proj.Native.greedy_online_chunk(frames, hyps, enc_is_raw=False)
            else:
                tokens = [list(s.Tokens) for s in active]
                n0 = [len(t) for t in tokens]
                tsl = [[] for _ in range(B)]
                decoder_out = proj.DecoderProj(hyps, B).decoder_out
                for t in range(frames.shape[1]):
                    logits = proj.JoinerProj(frames[:, t, :], decoder_out).Logits
                    ys = _argmax_hi_rows(logits)
                    emitted = False
                    for m in range(B):
                        if ys[m] != proj.Blank_id and ys[m] != proj.Unk_id and ys[m] != 1:     # ref :181
                            tokens[m].append(int(ys[m]))
                            tsl[m].append(t)
                            emitted = True
                    if emitted:
                        decoder_out = proj.DecoderProj(np.array([tk[-ctx:] for tk in tokens], np.int64), B).decoder_out
                toks = [tokens[m][n0[m]:] for m in range(B)]
                tss = tsl
                hyp_out = np.array([tk[-ctx:] for tk in tokens], np.int64)
            next_states = proj.unstack_states(enc.encoder_out_states)
            for m, s in enumerate(active):
                s.Tokens = list(s.Tokens) + toks[m]
                s.Hyp = np.asarray(hyp_out[m], np.int64).copy()          # ref :208
                s.Timestamps.extend(tss[m])                              # chunk-local t (ref :184)
                s.States = next_states[m] if next_states else s.States
                self._count_trailing(s, frames.shape[1], tss[m])
        except Exception as ex:
            raise Exception("Online recognition failed") from ex        # ref :215-218

    def ForwardBatchGreedySearchCTC(self, streams: List[OnlineStream]):  # ref :220-319
        if len(streams) == 0:
            return
        inputs, active = self._collect(streams)
        if not inputs:
            return
        proj = self._onlineProj
        B = len(inputs)
        try:
            states = proj.stack_states([s.States for s in active])
            enc = proj.EncoderProj(inputs, B, states)
            V = len(self._tokens) if self._tokens is not None else proj.CustomMetadata.Vocab_size
            logp = enc.encoder_out.reshape(B, -1, V)
            fo = np.array([s.FrameOffset for s in active], np.int32)
            toks, tss, _, _ = proj.Native.ctc_greedy(logp, blank=proj.Blank_id, frame_offset=fo)
            next_states = proj.unstack_states(enc.encoder_out_states)
            for m, s in enumerate(active):
                # Q10: FrameOffset / NumTrailingBlank are NOT written back, prev_id resets per chunk (ref :276, :300-313)
                s.Tokens = list(s.Tokens) + toks[m]
                s.Timestamps = list(s.Timestamps) + tss[m]
                s.States = next_states[m] if next_states else s.States
        except Exception as ex:
            raise Exception("Online recognition failed") from ex

    # -- streaming modified_beam_search (ours; what maxActivePaths was meant for) ------------------------------------------------
    def ForwardBatchModifiedBeamSearch(self, streams: List[OnlineStream]):
        if len(streams) == 0:
            return
        inputs, active = self._collect(streams)
        if not inputs:
            return
        proj = self._onlineProj
        B = len(inputs)
        try:
            states = proj.stack_states([s.States for s in active])
            enc = proj.EncoderProj(inputs, B, states)
            J = proj.CustomMetadata.Joiner_dim
            frames = enc.encoder_out.reshape(B, -1, J)
            slots = [s.BeamSlot for s in active]
            toks, tss, score, hyp_out = proj.Native.modified_beam_search_online_chunk(frames, slots, cap=self._cap, enc_is_raw=False)
            next_states = proj.unstack_states(enc.encoder_out_states)
            blank = proj.Blank_id
            for m, s in enumerate(active):
                # the best hypothesis may change retroactively: Tokens / Timestamps are REPLACED, not appended to
                s.Tokens = [blank, blank] + toks[m]                      # seed of ref OnlineStream.cs:45
                s.Timestamps = list(tss[m])                              # frame index since the stream began
                s.Hyp = np.asarray(hyp_out[m], np.int64).copy()
                s.States = next_states[m] if next_states else s.States
                s.NumProcessedFrames += frames.shape[1]
                s.NumTrailingBlank = s.NumProcessedFrames - 1 - tss[m][-1] if tss[m] else s.NumProcessedFrames
            self.LastScores = [float(x) for x in score]
        except Exception as ex:
            raise Exception("Online recognition failed") from ex

    def Dispose(self):
        self._onlineProj.Dispose()
