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
    }
}
