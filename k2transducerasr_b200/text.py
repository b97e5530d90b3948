"""Token ids -> text: the wire / format step after the search, host side, cold.

Restates, with citations, what the reference does with stream.Tokens once a Forward* delegate returned:
  DecodeMulti   ref OfflineRecognizer.cs:432-467 (offline: skips -1), OnlineRecognizer.cs:321-351 (online)
  CheckText     ref OfflineRecognizer.cs:492-546   <0xNN> byte-fallback runs -> UTF-8; otherwise byte-level-BPE decoding
  HexToStr      ref OfflineRecognizer.cs:553-575
  ByteDecode / SmartByteDecode   ref Utils/ByteDataHelper.cs:331-397 (a port of fairseq / icefall byte_utils)
  NbestHotwords ref Utils/HotwordsHelper.cs:8-57 (never called by the reference; kept for the same callers' benefit)
The search kernels hand over int64 ids; nothing here touches the GPU.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence

# Printable stand-ins of the 256 byte values (ref Utils/ByteDataHelper.cs:27-285, given there as a literal list): bytes 0..31 map to
# U+0100..U+011F, printable ASCII 32..126 to itself, 127..255 to Latin Extended characters with a few code points skipped.
_BYTE_RANGES = ((256, 287), (32, 126), (288, 305), (308, 318), (321, 328), (330, 382), (384, 422))
PRINTABLE_BASE_CHARS: List[int] = [c for lo, hi in _BYTE_RANGES for c in range(lo, hi + 1)]
assert len(PRINTABLE_BASE_CHARS) == 256
BPE_UNK = chr(8263)                                                      # ref :24
BYTE_TO_BCHAR: Dict[int, str] = {b: chr(PRINTABLE_BASE_CHARS[b]) for b in range(256)}     # ref :292-299
BCHAR_TO_BYTE: Dict[str, int] = {v: k for k, v in BYTE_TO_BCHAR.items()}                  # ref :300-305
BCHAR_TO_BYTE[BPE_UNK] = 32

_WS = re.compile(r"\s+")
_TAG = re.compile(r"<(\w+)>")                                            # ref OfflineRecognizer.cs:494
_CJK_ALL = re.compile(r"^[一-龥]+$")                             # ref :480
_SKIP = ("<blk>", "<sos/eos>", "<unk>")


def byte_encode(x: str) -> str:
    """ref ByteDataHelper.cs:313-324."""
    return "".join(BYTE_TO_BCHAR[b] for b in _WS.sub(" ", x).encode("utf-8"))


def byte_decode(x: str) -> str:
    """ref ByteDataHelper.cs:331-346: a character outside the table leaves the input unchanged (the catch branch); invalid UTF-8 is
    NOT an error in .NET's Encoding.UTF8.GetString - it yields U+FFFD, as errors='replace' does here."""
    try:
        return bytes(BCHAR_TO_BYTE[c] for c in x).decode("utf-8", errors="replace")
    except KeyError:
        return x


def smart_byte_decode(x: str) -> str:
    """ref ByteDataHelper.cs:353-397: plain decoding first; only when that yields the empty string, the longest-valid-pieces
    dynamic programme (pieces of up to 4 characters)."""
    out = byte_decode(x)
    if out == "":
        n = len(x)
        f = [0] * (n + 1)
        pt = [0] * (n + 1)
        for i in range(1, n + 1):
            f[i], pt[i] = f[i - 1], i - 1
            for j in range(1, min(4, i) + 1):
                if f[i - j] + 1 > f[i] and len(byte_decode(x[i - j:i])) > 0:
                    f[i], pt[i] = f[i - j] + 1, i - j
        cur = n
        while cur > 0:
            if f[cur] == f[pt[cur]] + 1:
                out = byte_decode(x[pt[cur]:cur]) + out
            cur = pt[cur]
    return out


def hex_to_str(hex_digits: str) -> str:
    """ref OfflineRecognizer.cs:553-575: pairs of hex digits -> bytes -> UTF-8; an odd count is padded with "20"."""
    if len(hex_digits) % 2:
        hex_digits += "20"
    try:
        raw = bytes(int(hex_digits[2 * i:2 * i + 2], 16) for i in range(len(hex_digits) // 2))   # (a digit left over is dropped)
    except ValueError as ex:
        raise ValueError("hex is not a valid hex number!") from ex
    return raw.decode("utf-8", errors="replace")


def check_text(text: str) -> str:
    """ref OfflineRecognizer.cs:492-546. Without any <...> tag the text is stripped of spaces and byte-BPE decoded; with tags,
    every run of adjacent six-character tags (<0xE4><0xBD><0xA0>) is replaced by the UTF-8 string of its bytes."""
    matches = list(_TAG.finditer(text))
    if not matches:
        return smart_byte_decode(text.replace(" ", ""))
    runs: List[str] = []
    cur, last = "", -1
    for k, m in enumerate(matches):
        if last == -1 or m.start() - last == 6:
            cur += m.group(0)
        else:
            runs.append(cur)
            cur = m.group(0)
        if k == len(matches) - 1:
            runs.append(cur)
        last = m.start()
    for run in runs:
        text = text.replace(run, hex_to_str(run.replace("<0x", "").replace(">", "")))
    return text


def decode_multi(token_ids: Sequence[int], symbols: Optional[Sequence[str]], online: bool = False):
    """One stream of DecodeMulti. Stops at the first id 2 (ref :443-446), offline skips -1 (ref :447-450; the online loop has no
    such test, its lists never hold -1), drops <blk> / <sos/eos> / <unk>, joins the first space-separated field of each tokens.txt
    line, maps U+2581 to a space, runs check_text and lower-cases. Returns (text, kept symbols)."""
    if symbols is None:
        return "", []
    kept: List[str] = []
    for t in token_ids:
        t = int(t)
        if t == 2:
            break
        if t == -1 and not online:
            continue
        s = symbols[t].split(" ")[0]
        if s in _SKIP:
            continue
        kept.append(s)                      # (the IsChinese branch of ref :453-460 appends the same string either way)
    return check_text("".join(kept).replace("▁", " ")).lower(), kept


def nbest_hotwords(token_nums: List[List[int]], token_nums_nbest: List[List[Sequence[int]]],
                   hotwords_list: Sequence[Sequence[int]]) -> List[List[int]]:
    """ref Utils/HotwordsHelper.cs:8-57 (dead code there: no caller). For every hot word (a token-id sequence) and every stream i:
    walk the per-frame n-best lists; while frame j's n-best holds the next token of the hot word, remember (j, token); when the
    whole hot word has been seen on consecutive frames, the frame after it triggers the substitution of those positions of
    token_nums[i] by the hot word's tokens; a frame that misses resets the match."""
    for hot in hotwords_list:
        for i, item_nbest in enumerate(token_nums_nbest):
            p = 0
            pos: List[int] = []
            words: List[int] = []
            for j, item in enumerate(item_nbest):
                if p < len(hot):
                    if hot[p] in item:
                        pos.append(j)
                        words.append(hot[p])
                    else:
                        pos, words, p = [], [], 0     # NB ref :34-35 resets `position` only; position_words keeps growing there,
                        continue                      # which mis-aligns later substitutions - we reset both
                    p += 1
                else:
                    for x, j_sub in enumerate(pos):
                        token_nums[i][j_sub] = words[x]
                    pos, words, p = [], [], 0
    return token_nums
