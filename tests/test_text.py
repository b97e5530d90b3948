"""CPU: token ids -> text (k2transducerasr_b200/text.py), the restatement of DecodeMulti / CheckText / HexToStr / ByteDataHelper
(ref OfflineRecognizer.cs:432-575, OnlineRecognizer.cs:321-351, Utils/ByteDataHelper.cs:313-397, Utils/HotwordsHelper.cs:8-57)."""
import pytest

from k2transducerasr_b200 import text as T


def test_byte_table_is_a_bijection_and_round_trips():
    assert len(set(T.PRINTABLE_BASE_CHARS)) == 256
    assert T.PRINTABLE_BASE_CHARS[:3] == [256, 257, 258] and T.PRINTABLE_BASE_CHARS[32] == 32 and T.PRINTABLE_BASE_CHARS[126] == 126
    assert T.PRINTABLE_BASE_CHARS[127] == 288 and T.PRINTABLE_BASE_CHARS[255] == 422      # ends of the reference's literal list
    for s in ("hello world", "你好，世界", "naïve café", "a\tb  c"):
        enc = T.byte_encode(s)
        assert all(ord(c) in T.PRINTABLE_BASE_CHARS for c in enc)
        assert T.byte_decode(enc) == " ".join(s.split()) if "\t" in s else T.byte_decode(enc) == s
    assert T.byte_decode("汉") == "汉"                 # a character outside the table: the input comes back (catch branch)
    assert T.byte_decode(T.BPE_UNK) == " "


def test_smart_byte_decode_recovers_valid_pieces_only_when_plain_decoding_is_empty():
    enc = T.byte_encode("你好")
    assert T.smart_byte_decode(enc) == "你好"
    assert T.smart_byte_decode("") == ""
    # .NET's GetString replaces invalid UTF-8 instead of failing, so a truncated sequence decodes to U+FFFD, not to ""
    assert "�" in T.smart_byte_decode(enc[:-1])


def test_hex_runs_and_check_text():
    assert T.hex_to_str("E4BDA0") == "你"
    assert T.hex_to_str("4") == "B"                     # odd length: "20" is appended, then Length / 2 pairs are read (ref :557-565)
    with pytest.raises(ValueError):
        T.hex_to_str("ZZ")
    # adjacent six-character tags form one run; a gap starts a new run
    assert T.check_text("<0xE4><0xBD><0xA0> ok <0xE5><0xA5><0xBD>") == "你 ok 好"
    assert T.check_text(" HELLO WORLD") == "HELLOWORLD"     # tag-free text loses its spaces before byte-BPE decoding (ref :497-499)
    assert T.check_text("你 好") == "你好"


def test_decode_multi_offline_and_online():
    syms = ["<blk> 0", "<sos/eos> 1", "<unk> 2", "▁HE 3", "LLO 4", "<0xE4> 5", "<0xBD> 6", "<0xA0> 7", "好 8"]
    text, kept = T.decode_multi([-1, 0, 3, 4, 5, 6, 7, 8], syms)
    assert kept == ["▁HE", "LLO", "<0xE4>", "<0xBD>", "<0xA0>", "好"] and text == " hello你好"
    assert T.decode_multi([0, 0, 8, 2, 3], syms, online=True) == ("好", ["好"])       # stops at id 2
    assert T.decode_multi([0, 1, 0], syms) == ("", [])
    assert T.decode_multi([3], None) == ("", [])
    # byte-level BPE vocabulary: the symbols are printable stand-ins of UTF-8 bytes
    bsyms = ["<blk> 0", "<sos/eos> 1", "<unk> 2"] + [f"{c} {i + 3}" for i, c in enumerate(T.byte_encode("你好"))]
    assert T.decode_multi(list(range(3, len(bsyms))), bsyms)[0] == "你好"


def test_nbest_hotwords_substitutes_a_complete_match():
    toks = [[5, 6, 7, 8, 9]]
    nbest = [[(5, 11), (20, 6), (21, 30), (8,), (9,)]]         # frames 1, 2 hold the hot word (20, 21) in their n-best
    out = T.nbest_hotwords([list(toks[0])], nbest, [(20, 21)])
    assert out == [[5, 20, 21, 8, 9]]
    assert T.nbest_hotwords([list(toks[0])], nbest, [(20, 99)]) == toks       # incomplete match: untouched
