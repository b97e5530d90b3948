// P/Invoke surface of libk2b200.so (include/k2b200.h). Blittable arguments only.
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no .NET toolchain. The same entry points are
// exercised through ctypes by k2transducerasr_b200/_native.py and tests/.
using System;
using System.Runtime.InteropServices;

namespace K2TransducerAsr.B200
{
    [StructLayout(LayoutKind.Sequential)]
    internal struct K2bConfig
    {
        public int struct_size, device, vocab_size, joiner_dim, decoder_dim, encoder_dim, context_size;
        public int blank_id, sos_eos_id, unk_id, max_streams, max_frames, max_beam, neg_id_mode, precision, reserved;
    }

    internal static class NativeMethods
    {
        private const string Lib = "k2b200";   // libk2b200.so / k2b200.dll next to the assembly

        public const int K2B_OK = 0;
        public const int GREEDY_SINGLE = 0, GREEDY_BATCH_COMPAT = 1, GREEDY_PER_STREAM = 2;
        public const int PREC_FP32 = 0, PREC_BF16X3 = 1, PREC_BF16 = 2;

        [DllImport(Lib)] public static extern int k2b_abi_version();
        [DllImport(Lib)] public static extern int k2b_create(ref K2bConfig cfg, out IntPtr handle);
        [DllImport(Lib)] public static extern int k2b_destroy(IntPtr h);
        [DllImport(Lib)] public static extern IntPtr k2b_last_error(IntPtr h);
        [DllImport(Lib)] public static extern int k2b_load_weights(IntPtr h, float[] emb, float[] conv_w, float[] dec_proj_w,
            float[] dec_proj_b, float[]? enc_proj_w, float[]? enc_proj_b, float[] out_w, float[] out_b);
        [DllImport(Lib)] public static extern int k2b_set_precision(IntPtr h, int precision);

        // fine-grained: 1:1 with IOfflineProj / IOnlineProj
        [DllImport(Lib)] public static extern int k2b_decoder_proj(IntPtr h, long[]? y, int n, [Out] float[] outp);
        [DllImport(Lib)] public static extern int k2b_joiner_proj(IntPtr h, float[] enc, float[] dec, int n, [Out] float[] logits);
        [DllImport(Lib)] public static extern int k2b_encoder_proj(IntPtr h, float[] raw, int n, [Out] float[] outp);

        // fused: 1:1 with the Forward* delegates
        // per-stream frame counts (EncoderOutputEntity.encoder_out_lens, which the stock search loops never read) for the next
        // k2b_greedy_offline (SINGLE / PER_STREAM) or k2b_modified_beam_search call
        [DllImport(Lib)] public static extern int k2b_set_encoder_out_lens(IntPtr h, long[]? lens, int B);
        [DllImport(Lib)] public static extern int k2b_greedy_offline(IntPtr h, float[] enc, int enc_is_raw, int B, int T, int mode,
            [Out] long[] tokens, [Out] int[] ts, [Out] int[] n_out, int cap);
        [DllImport(Lib)] public static extern int k2b_greedy_online_chunk(IntPtr h, float[] enc, int enc_is_raw, int B, int Tc,
            [In, Out] long[] hyp_inout, [Out] long[] tokens, [Out] int[] ts, [Out] int[] n_out, int cap);
        [DllImport(Lib)] public static extern int k2b_modified_beam_search(IntPtr h, float[] enc, int enc_is_raw, int B, int T, int K,
            [Out] long[] tokens, [Out] int[] ts, [Out] int[] n_out, [Out] float[] score, int cap);
        [DllImport(Lib)] public static extern int k2b_ctc_greedy(IntPtr h, float[] logp, int B, int T, int V, int blank,
            int[]? frame_offset, [In, Out] long[]? prev_inout, [Out] long[] tokens, [Out] int[] ts, [Out] int[] n_out,
            [In, Out] int[]? trailing_blank_inout, int cap);

        // on-device streaming state (replaces the Array.Copy loops of stack_states / unstack_states, ref OnlineProjOfZipformer2.cs:144-489):
        // one slot per OnlineStream; `stacked` is a DEVICE pointer of k2b_state_pool_stacked_floats(h, B) floats that the encoder
        // session binds as its cache inputs / outputs (ORT IOBinding on the CUDA EP), so the caches never cross PCIe
        [DllImport(Lib)] public static extern int k2b_state_pool_create(IntPtr h, int[] item_len, int n_tensors, int max_streams);
        [DllImport(Lib)] public static extern long k2b_state_pool_stacked_floats(IntPtr h, int B);
        [DllImport(Lib)] public static extern int k2b_state_pool_put(IntPtr h, int slot, float[] state);
        [DllImport(Lib)] public static extern int k2b_state_pool_get(IntPtr h, int slot, [Out] float[] state);
        [DllImport(Lib)] public static extern int k2b_stack_states(IntPtr h, int[] slots, int B, int[] axis_len, IntPtr stackedDev);
        [DllImport(Lib)] public static extern int k2b_unstack_states(IntPtr h, int[] slots, int B, int[] axis_len, IntPtr stackedDev);

        // device-pointer variants (frames already on the GPU, e.g. the encoder session's CUDA output): enqueue only, then k2b_sync
        [DllImport(Lib)] public static extern int k2b_set_stream(IntPtr h, IntPtr cudaStream);
        [DllImport(Lib)] public static extern int k2b_sync(IntPtr h);
        [DllImport(Lib)] public static extern int k2b_greedy_offline_dev(IntPtr h, IntPtr enc, int enc_is_raw, int B, int T, int mode,
            IntPtr tokens, IntPtr ts, IntPtr n_out, int cap);
        [DllImport(Lib)] public static extern int k2b_greedy_online_chunk_dev(IntPtr h, IntPtr enc, int enc_is_raw, int B, int Tc,
            IntPtr hyp_inout, IntPtr tokens, IntPtr ts, IntPtr n_out, int cap);
        [DllImport(Lib)] public static extern int k2b_modified_beam_search_dev(IntPtr h, IntPtr enc, int enc_is_raw, int B, int T, int K,
            IntPtr tokens, IntPtr ts, IntPtr n_out, IntPtr score, int cap);

        // round 2 of the ABI (K2B_ABI_VERSION 2)
        [DllImport(Lib)] public static extern int k2b_set_option(IntPtr h, [MarshalAs(UnmanagedType.LPUTF8Str)] string name, int value);
        [DllImport(Lib)] public static extern int k2b_get_stat(IntPtr h, [MarshalAs(UnmanagedType.LPUTF8Str)] string name, out double value);
        // page-locked host buffers: a managed float[] is pageable (staged, synchronous copies); frames / results that should move
        // at the PCIe rate live in memory from k2b_host_alloc (wrap it in a Span<float>) or in a pinned GCHandle registered here
        [DllImport(Lib)] public static extern int k2b_host_alloc(out IntPtr p, long bytes);
        [DllImport(Lib)] public static extern int k2b_host_free(IntPtr p);
        [DllImport(Lib)] public static extern int k2b_host_register(IntPtr p, long bytes);
        [DllImport(Lib)] public static extern int k2b_host_unregister(IntPtr p);
        // pointer-typed overloads of the fused calls for such buffers
        [DllImport(Lib, EntryPoint = "k2b_modified_beam_search")] public static extern int k2b_modified_beam_search_p(IntPtr h, IntPtr enc,
            int enc_is_raw, int B, int T, int K, IntPtr tokens, IntPtr ts, IntPtr n_out, IntPtr score, int cap);
        [DllImport(Lib, EntryPoint = "k2b_greedy_offline")] public static extern int k2b_greedy_offline_p(IntPtr h, IntPtr enc, int enc_is_raw,
            int B, int T, int mode, IntPtr tokens, IntPtr ts, IntPtr n_out, int cap);
        // streaming modified_beam_search: what decodingMethod / maxActivePaths of OnlineRecognizer (ref OnlineRecognizer.cs:18-19) select
        [DllImport(Lib)] public static extern int k2b_beam_pool_create(IntPtr h, int max_streams, int K, int max_frames);
        [DllImport(Lib)] public static extern int k2b_beam_pool_reset(IntPtr h, int slot, long[]? hyp);
        [DllImport(Lib)] public static extern int k2b_modified_beam_search_online_chunk(IntPtr h, float[] enc, int enc_is_raw, int B, int Tc,
            int[] slots, [Out] long[]? hyp_out, [Out] long[] tokens, [Out] int[] ts, [Out] int[] n_out, [Out] float[] score, int cap);
        [DllImport(Lib)] public static extern int k2b_modified_beam_search_online_chunk_dev(IntPtr h, IntPtr enc, int enc_is_raw, int B, int Tc,
            int[] slots, IntPtr hyp_out, IntPtr tokens, IntPtr ts, IntPtr n_out, IntPtr score, int cap);
        [DllImport(Lib)] public static extern int k2b_ctc_greedy_dev(IntPtr h, IntPtr logp, int B, int T, int V, int blank, IntPtr frame_offset,
            IntPtr prev_inout, IntPtr tokens, IntPtr ts, IntPtr n_out, IntPtr trailing_blank_inout, int cap);
        // multi-GPU reporting: one all-gather of the ranks' results (libnccl is loaded on first use)
        [DllImport(Lib)] public static extern int k2b_nccl_unique_id([Out] byte[] id128);
        [DllImport(Lib)] public static extern int k2b_nccl_init(IntPtr h, byte[] id128, int rank, int nranks);
        [DllImport(Lib)] public static extern int k2b_gather_results_nccl(IntPtr h, IntPtr tokens, IntPtr ts, IntPtr n, IntPtr score, int B, int cap,
            IntPtr all_tokens, IntPtr all_ts, IntPtr all_n, IntPtr all_score);
        [DllImport(Lib)] public static extern int k2b_gather_join(IntPtr h);
        // contextual biasing (hot words): dense automaton over token ids, built on the host (hotwords.py shows the construction)
        [DllImport(Lib)] public static extern int k2b_set_context_graph(IntPtr h, int[]? next, float[]? delta, float[]? residual, int n_states);
        [DllImport(Lib)] public static extern int k2b_debug_backpointers(IntPtr h, [Out] int[] outp, int B, int T, int K);

        internal static void Check(IntPtr h, int status, string what)
        {
            if (status == K2B_OK) return;
            string msg = Marshal.PtrToStringUTF8(k2b_last_error(h)) ?? "";
            throw new Exception($"{what} failed (libk2b200 status {status}): {msg}");
        }
    }
}
