// The fused replacements of the Forward* delegates (ref OfflineRecognizer.cs:10-11, :54-68; OnlineRecognizer.cs:46-57):
// one native call per GetResult(s). Each writes back exactly what the reference loops leave in the streams.
// NOT COMPILED HERE (no .NET toolchain); mirrored and tested by k2transducerasr_b200/recognizer.py.
using System;
using System.Collections.Generic;
using System.Linq;

namespace K2TransducerAsr.B200
{
    internal static class FusedSearch
    {
        // replaces ForwardBatchGreedySearch (ref OfflineRecognizer.cs:189-303)
        public static void ForwardBatchGreedySearch(OfflineProjOfB200 proj, List<OfflineStream> streams)
        {
            try
            {
                int B = streams.Count, J = proj.CustomMetadata.Joiner_dim;
                var enc = proj.EncoderProj(streams.Select(s => s.OfflineInputEntity).ToList(), B).encoder_out!;
                int T = enc.Length / J / B;
                if (T == 0) return;                                       // Q13
                var tokens = new long[B * T]; var ts = new int[B * T]; var n = new int[B];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_greedy_offline(proj.Native, enc, 0, B, T,
                    NativeMethods.GREEDY_BATCH_COMPAT, tokens, ts, n, T), "k2b_greedy_offline");
                for (int m = 0; m < B; m++)
                {
                    var tk = Enumerable.Repeat((long)proj.Blank_id, 2 * B).ToList();   // Q5 seed (ref :250-267)
                    var tt = Enumerable.Repeat(0, 2 * B).ToList();
                    for (int i = 0; i < n[m]; i++) { tk.Add(tokens[m * T + i]); tt.Add(ts[m * T + i]); }
                    streams[m].Tokens = tk;
                    streams[m].Timestamps.AddRange(tt);
                    streams[m].RemoveSamples();
                }
            }
            catch (Exception ex) { throw new Exception("Offline recognition failed", ex); }   // ref :299-302
        }

        // modified_beam_search (absent from the reference; selected by decodingMethod == "modified_beam_search")
        public static void ForwardBatchModifiedBeamSearch(OfflineProjOfB200 proj, List<OfflineStream> streams, int maxActivePaths)
        {
            try
            {
                int B = streams.Count, J = proj.CustomMetadata.Joiner_dim;
                var enc = proj.EncoderProj(streams.Select(s => s.OfflineInputEntity).ToList(), B).encoder_out!;
                int T = enc.Length / J / B;
                var tokens = new long[B * Math.Max(T, 1)]; var ts = new int[B * Math.Max(T, 1)]; var n = new int[B]; var score = new float[B];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_modified_beam_search(proj.Native, enc, 0, B, T, maxActivePaths,
                    tokens, ts, n, score, Math.Max(T, 1)), "k2b_modified_beam_search");
                for (int m = 0; m < B; m++)
                {
                    var tk = new List<long> { -1, proj.Blank_id };
                    for (int i = 0; i < n[m]; i++) { tk.Add(tokens[m * T + i]); streams[m].Timestamps.Add(ts[m * T + i]); }
                    streams[m].Tokens = tk;
                    streams[m].RemoveSamples();
                }
            }
            catch (Exception ex) { throw new Exception("Offline recognition failed", ex); }
        }

        // replaces ForwardBatchGreedySearchCTC (ref OfflineRecognizer.cs:365-430)
        public static void ForwardBatchGreedySearchCTC(IntPtr h, float[] logProbs, int vocab, int blank, List<OfflineStream> streams)
        {
            try
            {
                int B = streams.Count, T = logProbs.Length / B / vocab;
                var tokens = new long[B * Math.Max(T, 1)]; var ts = new int[B * Math.Max(T, 1)]; var n = new int[B];
                var fo = streams.Select(s => s.FrameOffset).ToArray();
                var tb = streams.Select(s => s.NumTrailingBlank).ToArray();
                NativeMethods.Check(h, NativeMethods.k2b_ctc_greedy(h, logProbs, B, T, vocab, blank, fo, null, tokens, ts, n, tb, Math.Max(T, 1)), "k2b_ctc_greedy");
                for (int m = 0; m < B; m++)
                {
                    for (int i = 0; i < n[m]; i++) { streams[m].Tokens.Add(tokens[m * T + i]); streams[m].Timestamps.Add(ts[m * T + i]); }
                    streams[m].NumTrailingBlank = tb[m];
                    streams[m].RemoveSamples();
                }
            }
            catch (Exception ex) { throw new Exception("Speech recognition failed", ex); }   // ref :426-429
        }

        // replaces ForwardGreedySearch, single stream (ref OfflineRecognizer.cs:93-187): Tokens = {-1, blank} + emitted (ref :115-117)
        public static void ForwardGreedySearch(OfflineProjOfB200 proj, OfflineStream stream)
        {
            try
            {
                int J = proj.CustomMetadata.Joiner_dim;
                var enc = proj.EncoderProj(new List<K2TransducerAsr.Model.OfflineInputEntity> { stream.OfflineInputEntity }, 1).encoder_out!;
                int T = enc.Length / J, cap = Math.Max(T, 1);
                var tokens = new long[cap]; var ts = new int[cap]; var n = new int[1];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_greedy_offline(proj.Native, enc, 0, 1, T, NativeMethods.GREEDY_SINGLE, tokens, ts, n, cap), "k2b_greedy_offline");
                var tk = new List<long> { -1, proj.Blank_id };
                for (int i = 0; i < n[0]; i++) { tk.Add(tokens[i]); stream.Timestamps.Add(ts[i]); }
                stream.Tokens = tk;
            }
            catch (Exception ex) { throw new Exception("Offline recognition failed", ex); }   // ref :183-186
        }

        // replaces ForwardGreedySearchCTC, single stream (ref OfflineRecognizer.cs:305-363); the reference swallows every exception here (Q11)
        public static void ForwardGreedySearchCTC(IntPtr h, float[] logProbs, int vocab, int blank, OfflineStream stream)
        {
            try
            {
                int T = logProbs.Length / vocab, cap = Math.Max(T, 1);
                var tokens = new long[cap]; var ts = new int[cap]; var n = new int[1]; var tb = new int[] { 0 };
                NativeMethods.Check(h, NativeMethods.k2b_ctc_greedy(h, logProbs, 1, T, vocab, blank, null, null, tokens, ts, n, tb, cap), "k2b_ctc_greedy");
                var tk = new List<long> { -1, blank };                                         // ref :318-320
                for (int i = 0; i < n[0]; i++) { tk.Add(tokens[i]); stream.Timestamps.Add(ts[i]); }
                stream.Tokens = tk; stream.NumTrailingBlank = tb[0];
            }
            catch (Exception) { }                                                              // ref :359-362
        }

        // ---- online: the delegates OnlineRecognizer's constructor binds (ref OnlineRecognizer.cs:46-57) ---------------------------
        // streams without a full chunk are removed from the caller's list, as the reference does (ref :97-120)
        private static (List<K2TransducerAsr.Model.OnlineInputEntity>, List<OnlineStream>) Collect(List<OnlineStream> streams)
        {
            var inputs = new List<K2TransducerAsr.Model.OnlineInputEntity>(); var active = new List<OnlineStream>(); var skipped = new List<OnlineStream>();
            foreach (var s in streams)
            {
                var chunk = s.GetDecodeChunk();
                if (chunk == null) { skipped.Add(s); continue; }
                inputs.Add(new K2TransducerAsr.Model.OnlineInputEntity { Speech = chunk, SpeechLength = chunk.Length });
                s.RemoveChunk();
                active.Add(s);
            }
            if (inputs.Count > 0) foreach (var s in skipped) streams.Remove(s);
            return (inputs, active);
        }

        // replaces OnlineRecognizer.ForwardBatchGreedySearch (ref OnlineRecognizer.cs:85-219): Hyp in / out (ref :109, :208), mask {blank, unk, 1}
        // (ref :181), chunk-local timestamps (ref :184)
        public static void ForwardBatchGreedySearchOnline(OnlineProjOfB200 proj, List<OnlineStream> streams)
        {
            if (streams.Count == 0) return;
            var (inputs, active) = Collect(streams);
            if (inputs.Count == 0) return;
            try
            {
                int B = inputs.Count, J = proj.CustomMetadata.Joiner_dim;
                var states = proj.stack_states(active.Select(s => s.States!).ToList());
                var encOut = proj.EncoderProj(inputs, B, states);
                var enc = encOut.encoder_out!;
                int Tc = enc.Length / J / B, cap = Math.Max(Tc, 1);
                var hyp = new long[2 * B];
                for (int m = 0; m < B; m++) Array.Copy(active[m].Hyp!, 0, hyp, 2 * m, 2);
                var tokens = new long[B * cap]; var ts = new int[B * cap]; var n = new int[B];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_greedy_online_chunk(proj.Native, enc, 0, B, Tc, hyp, tokens, ts, n, cap), "k2b_greedy_online_chunk");
                var next = proj.unstack_states(encOut.encoder_out_states!);
                for (int m = 0; m < B; m++)
                {
                    for (int i = 0; i < n[m]; i++) { active[m].Tokens!.Add(tokens[m * cap + i]); active[m].Timestamps!.Add(ts[m * cap + i]); }
                    active[m].Hyp = new long[] { hyp[2 * m], hyp[2 * m + 1] };
                    if (next.Count > m) active[m].States = next[m];
                    active[m].NumTrailingBlank = n[m] > 0 ? Tc - 1 - ts[m * cap + n[m] - 1] : active[m].NumTrailingBlank + Tc;   // what the endpoint rules read
                }
            }
            catch (Exception ex) { throw new Exception("Online recognition failed", ex); }   // ref :215-218
        }

        // streaming modified_beam_search: the hypotheses live in the stream's device slot (beamSlot[stream]); Tokens / Timestamps are REPLACED by
        // the best hypothesis so far (it may change retroactively), Hyp = its last two tokens. Selected by decodingMethod ==
        // "modified_beam_search" with maxActivePaths (ref OnlineRecognizer.cs:18-19: accepted and ignored by the reference).
        public static void ForwardBatchModifiedBeamSearchOnline(OnlineProjOfB200 proj, List<OnlineStream> streams, Dictionary<OnlineStream, int> beamSlot, int cap)
        {
            if (streams.Count == 0) return;
            var (inputs, active) = Collect(streams);
            if (inputs.Count == 0) return;
            try
            {
                int B = inputs.Count, J = proj.CustomMetadata.Joiner_dim;
                var states = proj.stack_states(active.Select(s => s.States!).ToList());
                var encOut = proj.EncoderProj(inputs, B, states);
                var enc = encOut.encoder_out!;
                int Tc = enc.Length / J / B;
                var slots = active.Select(s => beamSlot[s]).ToArray();
                var hyp = new long[2 * B]; var tokens = new long[B * cap]; var ts = new int[B * cap]; var n = new int[B]; var score = new float[B];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_modified_beam_search_online_chunk(proj.Native, enc, 0, B, Tc, slots, hyp, tokens, ts, n, score, cap),
                                    "k2b_modified_beam_search_online_chunk");
                var next = proj.unstack_states(encOut.encoder_out_states!);
                for (int m = 0; m < B; m++)
                {
                    var tk = new List<long> { proj.Blank_id, proj.Blank_id };                  // seed of ref OnlineStream.cs:45
                    var tt = new List<int>();
                    for (int i = 0; i < Math.Min(n[m], cap); i++) { tk.Add(tokens[m * cap + i]); tt.Add(ts[m * cap + i]); }
                    active[m].Tokens = tk; active[m].Timestamps = tt;
                    active[m].Hyp = new long[] { hyp[2 * m], hyp[2 * m + 1] };
                    if (next.Count > m) active[m].States = next[m];
                }
            }
            catch (Exception ex) { throw new Exception("Online recognition failed", ex); }
        }

        // replaces OnlineRecognizer.ForwardBatchGreedySearchCTC (ref OnlineRecognizer.cs:220-319) incl. Q10: prev_id resets per chunk (prev_inout
        // == null), FrameOffset / NumTrailingBlank are not written back (ref :276, :300-313)
        public static void ForwardBatchGreedySearchCTCOnline(OnlineProjOfB200 proj, string[] tokensTxt, List<OnlineStream> streams)
        {
            if (streams.Count == 0) return;
            var (inputs, active) = Collect(streams);
            if (inputs.Count == 0) return;
            try
            {
                int B = inputs.Count, V = tokensTxt.Length;                                    // ref :262: V = _tokens.Length
                var states = proj.stack_states(active.Select(s => s.States!).ToList());
                var encOut = proj.EncoderProj(inputs, B, states);
                var logp = encOut.encoder_out!;
                int T = logp.Length / B / V, cap = Math.Max(T, 1);
                var fo = active.Select(s => s.FrameOffset).ToArray();
                var tokens = new long[B * cap]; var ts = new int[B * cap]; var n = new int[B];
                NativeMethods.Check(proj.Native, NativeMethods.k2b_ctc_greedy(proj.Native, logp, B, T, V, proj.Blank_id, fo, null, tokens, ts, n, null, cap), "k2b_ctc_greedy");
                var next = proj.unstack_states(encOut.encoder_out_states!);
                for (int m = 0; m < B; m++)
                {
                    for (int i = 0; i < n[m]; i++) { active[m].Tokens!.Add(tokens[m * cap + i]); active[m].Timestamps!.Add(ts[m * cap + i]); }
                    if (next.Count > m) active[m].States = next[m];
                }
            }
            catch (Exception ex) { throw new Exception("Online recognition failed", ex); }
        }
    }
}
