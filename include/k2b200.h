/*
 * k2b200.h — C ABI of libk2b200.so: the transducer-search hot path of
 * manyeyes/K2TransducerAsr as hand-written sm_100a CUDA.
 *
 * Every entry point replaces one managed->native crossing (or one managed search loop) of the
 * reference. Citations "ref:" are file:line under the reference tree (K2TransducerAsr/...).
 *
 * Rules of the boundary (INTEGRATION.md shows the C# P/Invoke side):
 *   - blittable scalars and plain pointers only; no callbacks, no structs by value;
 *   - the caller owns every in/out buffer; the library owns all device memory;
 *   - every call returns an int status (K2B_OK == 0); k2b_last_error() gives the text;
 *   - a handle is bound to one CUDA device and one CUDA stream and is NOT thread-safe
 *     (the reference recognizers are unsynchronised too: ref OfflineRecognizer.cs:77-91);
 *   - token ids cross the boundary as int64 because the reference carries them as
 *     Int64[] / List<Int64> (ref IOfflineProj.cs:44, OfflineStream.cs:14); timestamps and
 *     counts are int32 (ref OfflineStream.cs:15).
 *   - entry points without a suffix take HOST pointers and are synchronous (H2D copy, kernels,
 *     D2H copy, stream sync inside the call). Entry points ending in _dev take DEVICE pointers,
 *     enqueue on the handle's stream and do NOT synchronise (call k2b_sync()).
 *   - there is no CPU fallback anywhere behind this header.
 */
#ifndef K2B200_H_
#define K2B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K2B_ABI_VERSION 2

#if defined(_WIN32)
#define K2B_API __declspec(dllexport)
#else
#define K2B_API __attribute__((visibility("default")))
#endif

/* status codes */
#define K2B_OK 0
#define K2B_ERR_INVALID 1   /* bad argument / shape / capacity              */
#define K2B_ERR_CUDA 2      /* a CUDA runtime or driver call failed (sticky) */
#define K2B_ERR_STATE 3     /* weights not loaded, handle poisoned, ...      */
#define K2B_ERR_UNSUPPORTED 4

/* greedy_offline modes */
#define K2B_GREEDY_SINGLE 0        /* ref OfflineRecognizer.cs:93-187 (B must be 1; 1000-symbol cap)   */
#define K2B_GREEDY_BATCH_COMPAT 1  /* ref OfflineRecognizer.cs:189-303 incl. the whole-batch decoder   */
                                   /* refresh on any emission                              */
#define K2B_GREEDY_PER_STREAM 2    /* every stream behaves as SINGLE regardless of its batch neighbours */

/* arithmetic of the three GEMMs (encoder_proj, decoder_proj, joiner output) */
#define K2B_PREC_FP32 0    /* CUDA-core FMA, fp32 operands                                   */
#define K2B_PREC_BF16X3 1  /* tcgen05, operands split hi+lo bf16, 3 MMAs, fp32 accumulate    */
#define K2B_PREC_BF16 2    /* tcgen05, operands rounded to bf16, fp32 accumulate             */

/* negative token id in the decoder context (ref OfflineRecognizer.cs:105 seeds {-1, blank}) */
#define K2B_NEGID_MASK 0   /* embedding row is zero (zipformer exports)        */
#define K2B_NEGID_WRAP 1   /* ONNX Gather semantics: -1 addresses row V-1      */

typedef struct k2b_handle k2b_handle;

/* Configuration = the ONNX custom-metadata contract the reference reads
 * (ref OfflineModel.cs:31-46: context_size, vocab_size, joiner_dim; ids ref OfflineModel.cs:18-20). */
typedef struct k2b_config {
  int32_t struct_size;   /* = sizeof(k2b_config); rejects ABI drift                   */
  int32_t device;        /* CUDA device ordinal                                       */
  int32_t vocab_size;    /* V                                                         */
  int32_t joiner_dim;    /* J  (the reference hard-codes 512 offline, Q8)             */
  int32_t decoder_dim;   /* D                                                         */
  int32_t encoder_dim;   /* E, raw encoder width fed to encoder_proj; 0 = none        */
  int32_t context_size;  /* must be 2 (ref OfflineRecognizer.cs:105-110, Q9)          */
  int32_t blank_id;      /* 0 */
  int32_t sos_eos_id;    /* 1 */
  int32_t unk_id;        /* 2 */
  int32_t max_streams;   /* largest B of any call                                     */
  int32_t max_frames;    /* largest T of any call                                     */
  int32_t max_beam;      /* largest K of any call (<= 8)                              */
  int32_t neg_id_mode;   /* K2B_NEGID_*                                               */
  int32_t precision;     /* K2B_PREC_*                                                */
  int32_t reserved;      /* 0                                                         */
} k2b_config;

/* ---- lifetime ------------------------------------------------------------------------------ */
K2B_API int32_t k2b_abi_version(void);
/* replaces: new OfflineModel/OnlineModel + proj construction (ref OfflineRecognizer.cs:30-53). */
K2B_API int32_t k2b_create(const k2b_config* cfg, k2b_handle** out);
/* replaces: IOfflineProj.Dispose / IOnlineProj.Dispose (ref IOfflineProj.cs:47). NULL is a no-op. */
K2B_API int32_t k2b_destroy(k2b_handle* h);
/* UTF-8 text of the last failure on this handle (h == NULL: last k2b_create failure). */
K2B_API const char* k2b_last_error(const k2b_handle* h);

/* Host fp32 row-major weights = the tensors inside decoder.onnx / joiner.onnx / encoder_proj
 * (ref _decoderSession/_joinerSession construction, OfflineModel.cs:84-118):
 *   emb [V,D]; conv_w [D,4,ctx] (Conv1d groups=D/4, no bias); dec_proj_w [J,D], dec_proj_b [J];
 *   enc_proj_w [J,E], enc_proj_b [J] (NULL when E == 0); out_w [V,J], out_b [V].            */
K2B_API int32_t k2b_load_weights(k2b_handle* h, const float* emb, const float* conv_w,
                         const float* dec_proj_w, const float* dec_proj_b,
                         const float* enc_proj_w, const float* enc_proj_b,
                         const float* out_w, const float* out_b);
/* switch K2B_PREC_* after creation (weights keep all derived copies). */
K2B_API int32_t k2b_set_precision(k2b_handle* h, int32_t precision);
/* Make the handle launch on a caller-owned cudaStream_t (e.g. torch's current stream); NULL
 * restores the handle's own stream. */
K2B_API int32_t k2b_set_stream(k2b_handle* h, void* cuda_stream);
/* Waits for everything enqueued on the handle's stream, then reports what the kernels flagged since the last check: an mbarrier /
 * counter time-out (K2B_ERR_STATE) or a Hyp with ids outside the vocabulary handed to a _dev entry point (K2B_ERR_INVALID).     */
K2B_API int32_t k2b_sync(k2b_handle* h);
/* Engine switches and search options by name (the K2B_* environment variables only seed them when the handle is created):
 *   "max_sym_per_frame" 1..16  symbols one frame may emit in k2b_greedy_offline SINGLE / PER_STREAM (ref OfflineRecognizer.cs:19,
 *                              :127-134; the reference fixes 1). Values > 1 run on the per-frame launches.
 *   "async_d2h" 0/1            host-pointer fused calls return once their copies are enqueued (give them pinned buffers, see
 *                              k2b_host_alloc); k2b_sync() completes them. Lets batch i's results leave while batch i+1 arrives.
 *   "pipe_chunks" 0..64        time chunks of the host-pointer beam search (0 = automatic)
 *   "inputs_complete" 0/1      1: the caller promises that frames handed to the _dev entry points are complete in device memory when
 *                              the call is made (an encoder that ran on another stream and was synchronised), not merely ordered on
 *                              the handle's stream. The device-pointer beam search (which projects raw frames in time chunks on
 *                              a side stream while one search launch polls per-chunk "projected" flags, "dev_chunks" 3) then does
 *                              not order that side stream behind earlier work of the handle's stream: the next call's frames are
 *                              projected under the current call's search.
 *   "async_gather" 0/1         k2b_gather_results_nccl on a side stream (see there)
 *   "copy_threads" -1..64      host threads that stage PAGEABLE inputs into the library's page-locked bounce buffers (-1 = a quarter of
 *                              the host's threads, 2..8; 0 = the calling thread copies). Page-locked inputs need none.
 *   "no_mega", "unfused_step", "greedy_persistent" (-1 auto / 0 / 1), "pair", "prof_which", "cluster_timing" (0 = off),
 *   "wh_tmem_kb" (-1 auto / 0), "single_greedy" (0 / 1), "dev_chunks" (-1 auto / 1 / 2 / 3), "dev_chunk_frames" (0 auto / 2..4096), "tagged_records" (0 / 1), "ctc_one_kernel" (-1 by input size / 0 / 1):
 *                              comparison switches between engines that must give identical results.        */
K2B_API int32_t k2b_set_option(k2b_handle* h, const char* name, int32_t value);
/* "decoder_table_bytes", "decoder_table_build_ms", "decoder_table_state" (0 not built, 1 built, -1 does not fit).               */
K2B_API int32_t k2b_get_stat(k2b_handle* h, const char* name, double* value);
/* number of this library's kernels launched on the handle since creation / last reset. */
K2B_API int64_t k2b_launch_count(const k2b_handle* h);
K2B_API int32_t k2b_reset_launch_count(k2b_handle* h);
/* When on, the dominant GEMM (joiner) of every step is bracketed with CUDA events on the launch
 * stream; k2b_profile_read returns launches seen and their summed device time. */
K2B_API int32_t k2b_profile_enable(k2b_handle* h, int32_t on);
K2B_API int32_t k2b_profile_read(k2b_handle* h, int64_t* n_launches, double* total_ms);

/* ---- fine-grained: 1:1 with the proj seam ---------------------------------------------------- */
/* replaces: IOfflineProj/IOnlineProj.DecoderProj (ref OfflineProjOfTransducer.cs:93-123 and the
 * four identical online copies). y [n,ctx] int64; y == NULL means n x {-1, blank}
 * (ref OfflineProjOfTransducer.cs:97-110). out [n,J].                                          */
K2B_API int32_t k2b_decoder_proj(k2b_handle* h, const int64_t* y, int32_t n, float* out);
K2B_API int32_t k2b_decoder_proj_dev(k2b_handle* h, const int64_t* y, int32_t n, float* out);
/* replaces: JoinerProj (ref OfflineProjOfTransducer.cs:125-152). enc, dec [n,J]; logits [n,V]. */
K2B_API int32_t k2b_joiner_proj(k2b_handle* h, const float* enc, const float* dec, int32_t n, float* logits);
K2B_API int32_t k2b_joiner_proj_dev(k2b_handle* h, const float* enc, const float* dec, int32_t n, float* logits);
/* replaces: the encoder_proj Linear that upstream folds into encoder.onnx, i.e. the tail of
 * EncoderProj (ref OfflineProjOfTransducer.cs:48-92). raw [n,E]; out [n,J].                    */
K2B_API int32_t k2b_encoder_proj(k2b_handle* h, const float* raw, int32_t n, float* out);
K2B_API int32_t k2b_encoder_proj_dev(k2b_handle* h, const float* raw, int32_t n, float* out);

/* ---- fused search loops: 1:1 with the Forward* delegates ------------------------------------- */
/* All fused calls write only the symbols APPENDED by this call (the reference's list seeds, e.g.
 * {-1, blank} or the 2B blanks of Q5, are re-created by the host shim). tokens/ts are [B,cap]
 * row-major, n_out [B]; cap >= T is required (at most one symbol per frame).
 * `enc` is [B,T,J] already projected when enc_is_raw == 0, or raw [B,T,E] (then encoder_proj
 * runs first, on device) when enc_is_raw != 0.                                                  */

/* Page-locked host buffers. Managed arrays handed over P/Invoke are pageable: their copies go through the driver's staging buffer
 * at a fraction of the PCIe rate and block the caller. Frame / result buffers taken from k2b_host_alloc (or registered in place
 * with k2b_host_register, e.g. a pinned GCHandle) are copied by DMA, asynchronously.                                              */
K2B_API int32_t k2b_host_alloc(void** out, int64_t bytes);
K2B_API int32_t k2b_host_free(void* p);
K2B_API int32_t k2b_host_register(void* p, int64_t bytes);
K2B_API int32_t k2b_host_unregister(void* p);

/* Ragged batches. The reference carries `encoder_out_lens` through its seam (ref Model/EncoderOutputEntity.cs:10-20,
 * OfflineProjOfTransducer.cs:84) but never consumes it: padded frames are decoded like speech (Q7). This call hands the
 * per-stream frame counts (HOST pointer, [B], each 0..T) to the NEXT k2b_greedy_offline[_dev] (modes SINGLE / PER_STREAM)
 * or k2b_modified_beam_search[_dev] call of this handle; stream b is then decoded over frames [0, lens[b]) only, its
 * hypotheses frozen afterwards. The setting is consumed by that call (an online chunk call in between discards it: the online
 * loops have no lengths). lens == NULL clears it. BATCH_COMPAT greedy (the reference's own loop, which couples the streams of a
 * batch) rejects it with K2B_ERR_UNSUPPORTED.                                                                          */
K2B_API int32_t k2b_set_encoder_out_lens(k2b_handle* h, const int64_t* lens, int32_t B);

/* replaces: ForwardGreedySearch / ForwardBatchGreedySearch (ref OfflineRecognizer.cs:93-303).
 * ts = frame index t (ref :164, :271).                                                          */
K2B_API int32_t k2b_greedy_offline(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T,
                           int32_t mode, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap);
K2B_API int32_t k2b_greedy_offline_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T,
                               int32_t mode, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap);

/* replaces: OnlineRecognizer.ForwardBatchGreedySearch for one chunk (ref OnlineRecognizer.cs:85-219).
 * hyp_inout [B,ctx] = OnlineStream.Hyp in, last ctx tokens out (ref :109, :208). Emission mask is
 * {blank, unk, 1} (ref :181); ts is chunk-local (ref :184).                                     */
K2B_API int32_t k2b_greedy_online_chunk(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap);
K2B_API int32_t k2b_greedy_online_chunk_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                    int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap);

/* modified_beam_search: ABSENT from the reference (only the dead maxActivePaths argument,
 * ref OnlineRecognizer.cs:19); semantics are icefall's, restated in oracle/k2_oracle.py (A.5).
 * Output = best hypothesis per stream by log_prob/len (len counts the 2 seed entries);
 * score [B] = its un-normalised log-prob.                                                       */
K2B_API int32_t k2b_modified_beam_search(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T,
                                 int32_t K, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score,
                                 int32_t cap);
K2B_API int32_t k2b_modified_beam_search_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T,
                                     int32_t K, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score,
                                     int32_t cap);

/* Streaming modified_beam_search: what `decodingMethod` / `maxActivePaths` of the online recognizer (ref OnlineRecognizer.cs:18-19,
 * accepted and ignored there, :46-57) select on this path. A stream's K live hypotheses and the back-pointer history of every
 * frame decoded so far live in a device slot (as the encoder caches do, k2b_state_pool_*); the slot takes the place of
 * OnlineStream.Hyp as the state carried between GetResults calls (ref OnlineStream.cs:44-55, OnlineRecognizer.cs:206-213).
 *   k2b_beam_pool_create   max_streams slots, beam K (1..8), max_frames frames of history per stream.
 *   k2b_beam_pool_reset    new utterance in `slot`: one hypothesis with context hyp[0..1] = OnlineStream.Hyp ({blank, blank},
 *                          ref OnlineStream.cs:44) or, hyp == NULL, the offline seed {-1, blank}.
 *   ..._online_chunk       decodes Tc more frames of the B streams whose slots are listed (HOST pointer, distinct) and returns, per
 *                          stream, the WHOLE best hypothesis so far (log_prob / len rule of k2b_modified_beam_search): tokens /
 *                          ts [B,cap] (ts = frame index since the reset; the first cap symbols when n_out exceeds cap), score, and
 *                          hyp_out [B,ctx] = its last ctx tokens (NULL: not wanted). Non-emitting ids: blank, unk and the literal 1
 *                          of the online loop (ref OnlineRecognizer.cs:181). Decoding an utterance chunk by chunk gives exactly
 *                          what k2b_modified_beam_search gives on the whole of it.                                               */
K2B_API int32_t k2b_beam_pool_create(k2b_handle* h, int32_t max_streams, int32_t K, int32_t max_frames);
K2B_API int32_t k2b_beam_pool_reset(k2b_handle* h, int32_t slot, const int64_t* hyp);
K2B_API int32_t k2b_modified_beam_search_online_chunk(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                                      const int32_t* slots, int64_t* hyp_out, int64_t* tokens, int32_t* ts,
                                                      int32_t* n_out, float* score, int32_t cap);
K2B_API int32_t k2b_modified_beam_search_online_chunk_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                                          const int32_t* slots, int64_t* hyp_out, int64_t* tokens, int32_t* ts,
                                                          int32_t* n_out, float* score, int32_t cap);

/* Contextual biasing (hot words) of modified_beam_search - SURVEY.md section 8f rank 4; the reference only holds a dead N-best
 * substitution stub (ref Utils/HotwordsHelper.cs:8-57, no caller). The hot words are compiled on the host into a dense automaton
 * over token ids (k2transducerasr_b200/hotwords.py, csharp: the same tables): HOST pointers next [S,V] int32, delta [S,V] float,
 * residual [S] float; state 0 is the root. When a hypothesis in state s is extended by token y (after the top-K selection, as
 * icefall does), delta[s,y] is added to its log-prob and it moves to next[s,y]; a match that breaks gives its boost back through a
 * negative delta; when the utterance ends in state s, residual[s] (the boost of a match still in progress) is subtracted before
 * the hypotheses are compared. Served by the cluster kernel (V <= 1024, beams 2 / 4 / 8, incl. the time-chunked and streaming
 * calls, where the state travels with the hypothesis) and by the per-frame merge (fp32 precision, or option "unfused_step");
 * other engines answer K2B_ERR_UNSUPPORTED while a graph is set. Greedy search is never biased. n_states == 0 clears it.      */
K2B_API int32_t k2b_set_context_graph(k2b_handle* h, const int32_t* next, const float* delta, const float* residual, int32_t n_states);

/* replaces: ForwardGreedySearchCTC / ForwardBatchGreedySearchCTC (ref OfflineRecognizer.cs:305-430)
 * and the online variant (ref OnlineRecognizer.cs:220-319). logp [B,T,V]; V is an argument because
 * the reference takes it from tokens.txt (ref :325). frame_offset [B] or NULL (= 0) is added to
 * ts (ref :349). prev_inout [B] or NULL: previous frame's argmax carried across calls (NULL = the
 * reference's per-call reset to -1, Q10). trailing_blank_inout [B] or NULL (ref :337-344).
 * Needs no weights.                                                                             */
K2B_API int32_t k2b_ctc_greedy(k2b_handle* h, const float* logp, int32_t B, int32_t T, int32_t V, int32_t blank,
                       const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts,
                       int32_t* n_out, int32_t* trailing_blank_inout, int32_t cap);
K2B_API int32_t k2b_ctc_greedy_dev(k2b_handle* h, const float* logp, int32_t B, int32_t T, int32_t V, int32_t blank,
                           const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts,
                           int32_t* n_out, int32_t* trailing_blank_inout, int32_t cap);

/* ---- on-device streaming state (the step either side of the online path) ---------------------- */
/* replaces: stack_states / unstack_states (ref OnlineProjOfZipformer2.cs:144-362, :363-489; OnlineProjOfZipformer.cs,
 * OnlineProjOfLstm.cs and OnlineProjOfConformer.cs hold the same loops), which re-lay the per-stream encoder caches into
 * batched tensors (batch on axis 1) with Array.Copy before every encoder call and back after it. Here a stream's caches
 * live concatenated in one slot of a device pool (OnlineStream.States, ref OnlineStream.cs:15) and the re-layout is one
 * launch. For cache tensor i with per-stream length item_len[i] and "axisnum" A = axis_len[i] (X = item_len/A):
 *     stacked_i[(x*B + n)*A + a] = slot(slots[n])_i[x*A + a],     stacked_i starts at float B * off_i of `stacked`,
 * off_i = sum of the lengths before i, each rounded up to 4 floats. axis_len is an argument of BOTH calls because the
 * reference does not use the same A in both directions for every tensor (cached_nonlin_attn: ref :250 vs :409).
 * item_len, axis_len, slots, state: HOST pointers; stacked: DEVICE pointer of k2b_state_pool_stacked_floats(h, B) floats. */
K2B_API int32_t k2b_state_pool_create(k2b_handle* h, const int32_t* item_len, int32_t n_tensors, int32_t max_streams);
K2B_API int64_t k2b_state_pool_stacked_floats(k2b_handle* h, int32_t B);
/* upload / download one stream's caches (plain concatenation of the tensors, e.g. from GetEncoderInitStates) */
K2B_API int32_t k2b_state_pool_put(k2b_handle* h, int32_t slot, const float* state);
K2B_API int32_t k2b_state_pool_get(k2b_handle* h, int32_t slot, float* state);
K2B_API int32_t k2b_stack_states(k2b_handle* h, const int32_t* slots, int32_t B, const int32_t* axis_len, float* stacked);
K2B_API int32_t k2b_unstack_states(k2b_handle* h, const int32_t* slots, int32_t B, const int32_t* axis_len, const float* stacked);

/* ---- multi-GPU: results of all ranks (reporting only; no collective runs inside a search) ------ */
/* One process per GPU, every rank decodes its own streams; the only exchange is one all-gather of the results over NVLink.
 * libnccl.so.2 is loaded on first use (no link-time dependency). Rank 0 obtains a 128-byte id with k2b_nccl_unique_id and
 * hands it to the other ranks by any means (bench.py: torch.distributed broadcast), every rank calls k2b_nccl_init once.
 * k2b_gather_results_nccl: DEVICE pointers; tokens / ts [B,cap], n / score [B] of this rank in, rank-major [nranks*B, ...] out;
 * enqueued on the handle's stream behind the search that produced the inputs, no host synchronisation. B and cap must be the same
 * on every rank. score / all_score may both be NULL.
 * With k2b_set_option("async_gather", 1) the gather runs on a side stream behind an event of the handle's stream, so the next
 * batch's search starts without waiting for it: the library orders its own later writes to result buffers behind it (the cluster
 * beam search only its back-trace, every other call as a whole); all_* are complete after k2b_gather_join (stream order, no host
 * synchronisation) or k2b_sync.                                                                                                */
K2B_API int32_t k2b_nccl_unique_id(void* id128);
K2B_API int32_t k2b_nccl_init(k2b_handle* h, const void* id128, int32_t rank, int32_t nranks);
K2B_API int32_t k2b_gather_results_nccl(k2b_handle* h, const int64_t* tokens, const int32_t* ts, const int32_t* n, const float* score,
                                        int32_t B, int32_t cap, int64_t* all_tokens, int32_t* all_ts, int32_t* all_n, float* all_score);
K2B_API int32_t k2b_gather_join(k2b_handle* h);

/* ---- diagnostics ---------------------------------------------------------------------------- */
/* Back-pointer rows of the LAST beam search of this handle, HOST pointer [B,T,K] int32: entry = (parent slot << 28) | (appended
 * token + 1), 0 for a dead slot - the whole history of every stream's beam, frame by frame. The parity tests compare it with
 * the oracle's beams to find the frame at which a stream first diverges (and so verify the 64-bit sequence-hash dedupe against
 * real token sequences).                                                                                                     */
K2B_API int32_t k2b_debug_backpointers(k2b_handle* h, int32_t* out, int32_t B, int32_t T, int32_t K);
/* Hardware self-tests of the tcgen05 / TMEM / bulk-TMA / cluster building blocks (HOST pointers).
 * k2b_selftest_umma: D[128,N] = A[128,K] * B[N,K]^T through tcgen05.mma; mode 0 = bf16 operands in
 * shared memory, 1 = split-bf16 x3 all in shared memory, 2 = x3 with the low part of A resident in
 * TMEM, 3 = bf16 with A resident in TMEM; use_tma != 0 stages A through one bulk TMA copy.          */
K2B_API int32_t k2b_selftest_umma(k2b_handle* h, const float* A, const float* B, int32_t N, int32_t K,
                                  int32_t mode, int32_t use_tma, float* D);
/* k2b_selftest_umma2: D[256,N] = bf16(A[256,K]) * bf16(B[N,K])^T through one CTA pair (tcgen05.mma.cta_group::2, B operand split
 * between the two CTAs, multicast commit); ts != 0 takes A from TMEM. N % 32 == 0, N <= 256; K % 64 == 0, K <= 256.               */
K2B_API int32_t k2b_selftest_umma2(k2b_handle* h, const float* A, const float* B, int32_t N, int32_t K, int32_t ts, float* D);
/* k2b_selftest_umma_bench: cycles to issue / complete reps*nkb*4 tcgen05.mma of one flavour
 * (0 SS N=32, 1 SS N=64, 2 TS N=32, 3 TS N=64, 4 SS M=64 N=32, 5 SS N=128, 6 TS N=128).             */
K2B_API int32_t k2b_selftest_umma_bench(k2b_handle* h, int32_t flavour, int32_t nkb, int32_t reps,
                                        int64_t* cycles2);
/* k2b_selftest_collectives: dependent-chain latency (cycles) of REDUX, a 5-level SHFL butterfly and BALLOT,
 * for one warp alone [0..2] and with 16 warps running the chain concurrently [3..5].                  */
K2B_API int32_t k2b_selftest_collectives(k2b_handle* h, int64_t* out6);
/* k2b_selftest_dsmem_bw: cycles for every CTA of a cluster to push bytes_per_peer to each peer with 16-byte
 * st.shared::cluster stores (two cluster barriers included).                                          */
K2B_API int32_t k2b_selftest_dsmem_bw(k2b_handle* h, int32_t csize, int32_t nclusters, int32_t bytes_per_peer,
                                      int64_t* cycles);
/* k2b_selftest_cluster: nclusters clusters of csize CTAs exchange data through distributed shared
 * memory; *bad = mismatching words, *ctas_done = CTAs that ran.                                     */
K2B_API int32_t k2b_selftest_cluster(k2b_handle* h, int32_t csize, int32_t nclusters, int32_t* bad,
                                     int32_t* ctas_done);

/* Cycle totals (out20: 20 int64 slots) of the phases of one frame step (0-7: build, MMA wait, TMEM read-out, reductions, exchange,
 * exchange wait, merge, end barrier; 8-14: merge sub-steps; 15-18: the four K-quarters of the build) summed over the last
 * cluster-kernel launch, CTA 0. The first call switches the collection on (the instrumented kernel instantiation then serves
 * every launch); k2b_set_option(h, "cluster_timing", 0) switches it off again.                                           */
K2B_API int32_t k2b_cluster_phase_cycles(k2b_handle* h, int64_t* out20);

/* Diagnostic clock64 timeline of the large-vocabulary beam search (joiner_tc.cu). The first call allocates the buffer and
 * switches the collection on; later calls copy out [64][148][8] int64 stamps and re-arm. Persistent kernel: CTA 0 writes
 * [frame][16] role stamps for the first 40 frames and [frame][8] merge-step stamps from element 640 on
 * (tools/timeline_mega.py); per-frame launches: [frame][sm][8] (tools/timeline_cfg4.py).                       */
K2B_API int32_t k2b_debug_timeline(k2b_handle* h, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* K2B200_H_ */
