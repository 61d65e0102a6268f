"""Tokenizer lookup: libqwen3.so's hash-table tokenizer against the reference's compiled src/tokenizer.c
(oracle/_ref/libqwen3_ref_tokenizer.so) and committed golden ids, on synthetic vocabularies with merges, duplicate strings,
"<..>" specials and unknown bytes. Host code only: runs without a GPU."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref", "libqwen3_ref_tokenizer.so")
GOLD = os.path.join(HERE, "golden", "tokenizer_golden.npz")


def write_vocab(path, seed=7, n_merges=3000, max_len=12):
    """256 byte-level entries (printable bytes as themselves, the rest as "<xx>") + merges made of existing entries, scores
    = -log(rank + 1) as the reference's exporter writes them (qwen3/tokenizer.py), with a few duplicate strings."""
    rng = np.random.default_rng(seed)
    toks = [bytes([i]) if 32 <= i < 127 else b"<%02x>" % i for i in range(256)]
    printable = [t for t in toks if len(t) == 1 and t not in (b"<", b">")]
    seen = set(toks)
    alphabet = [bytes([c]) for c in b"abcde "]
    pool = list(alphabet)
    while len(toks) < 256 + n_merges:
        a = pool[rng.integers(len(pool))]
        b = pool[rng.integers(len(pool))] if rng.random() < 0.7 else printable[rng.integers(len(printable))]
        new = a + b
        if len(new) > max_len:
            continue
        if new in seen and rng.random() > 0.02:  # a few duplicates on purpose: the lowest id must win
            continue
        seen.add(new)
        toks.append(new)
        pool.append(new)
    toks += [b"<|im_start|>", b"<|im_end|>"]
    with open(path + ".tokenizer", "wb") as f:
        f.write(struct.pack("<Iiii", 0x71746B6E, 2, len(toks), max(len(t) for t in toks)))
        f.write(struct.pack("<10i", len(toks) - 1, len(toks) - 2, *([-1] * 8)))
        for i, t in enumerate(toks):
            f.write(struct.pack("<fi", -float(np.log(i + 1.0)), len(t)))
            f.write(t)
    return toks


def texts(rng):
    out = [b"abc", b"a", b"<|im_start|>user\nabba cab<|im_end|>\n", b"<0a>x<zz>y<", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"\x01a\x02"]
    for n in (5, 40, 300, 1500):
        out.append(bytes(rng.choice(list(b"abcde  <>|"), size=n).tolist()))
    return out


class Tok:
    class _T(C.Structure):
        _fields_ = [("entries", C.c_void_p), ("special", C.c_int * 10), ("magic", C.c_int), ("version", C.c_int),
                    ("vocab_size", C.c_int), ("max_len", C.c_int)]

    def __init__(self, lib_path, prefix):
        L = self.L = C.CDLL(lib_path)
        L.tokenizer_create.restype = C.POINTER(self._T)
        L.tokenizer_create.argtypes = [C.c_char_p]
        L.tokenizer_free.argtypes = [C.POINTER(self._T)]
        L.tokenizer_token_to_id.argtypes = [C.POINTER(self._T), C.c_char_p]
        L.tokenizer_id_to_token.restype = C.c_char_p
        L.tokenizer_id_to_token.argtypes = [C.POINTER(self._T), C.c_int]
        L.tokenizer_encode.argtypes = [C.POINTER(self._T), C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.t = L.tokenizer_create(prefix.encode())
        assert self.t, lib_path

    def encode(self, text: bytes):
        buf = C.create_string_buffer(text)
        ids = (C.c_int * (len(text) + 1))()
        n = C.c_int(0)
        self.L.tokenizer_encode(self.t, buf, ids, C.byref(n))
        return list(ids[: n.value])

    def close(self):
        self.L.tokenizer_free(self.t)


@pytest.fixture(scope="module")
def ours_lib(pkg):
    pkg.build.build()
    return pkg.build.lib_path()


def test_tokenizer_matches_golden_and_reference(ours_lib, tmp_path):
    prefix = str(tmp_path / "vocab")
    toks = write_vocab(prefix)
    ours = Tok(ours_lib, prefix)
    assert ours.t.contents.vocab_size == len(toks) and ours.t.contents.special[0] == len(toks) - 1
    rng = np.random.default_rng(3)
    cases = texts(rng)
    got = [ours.encode(t) for t in cases]
    # decoding the ids gives the text back, minus the bytes the vocabulary does not have
    for t, ids in zip(cases, got):
        back = b"".join(ours.L.tokenizer_id_to_token(ours.t, i) for i in ids)
        assert back == bytes(c for c in t if 32 <= c < 127), t[:40]  # non-printable bytes exist only as "<xx>" entries
    gold = np.load(GOLD)  # ids the reference's own tokenizer.c produced for these texts (tests/golden/make_tokenizer_golden.py)
    for i, ids in enumerate(got):
        assert ids == gold[f"ids{i}"].tolist(), (i, cases[i][:40])
    # lookups: every string of the vocabulary (duplicates -> the lowest id), and strings that are not in it
    for i in list(range(0, len(toks), 37)) + [len(toks) - 1]:
        assert ours.L.tokenizer_token_to_id(ours.t, toks[i]) == toks.index(toks[i])
    assert ours.L.tokenizer_token_to_id(ours.t, b"no such token") == -1
    if os.path.exists(REF):  # the compiled reference, side by side (present wherever oracle/_ref was built)
        ref = Tok(REF, prefix)
        for t, ids in zip(cases, got):
            assert ref.encode(t) == ids, t[:40]
        for probe in (b"ab", b"abc", b"<0a>", b"zz", toks[-1], toks[300]):
            assert ref.L.tokenizer_token_to_id(ref.t, probe) == ours.L.tokenizer_token_to_id(ours.t, probe)
        ref.close()
    ours.close()


def test_tokenizer_create_rejects_bad_files(ours_lib, tmp_path):
    L = Tok.__new__(Tok)
    lib = C.CDLL(ours_lib)
    lib.tokenizer_create.restype = C.c_void_p
    lib.tokenizer_create.argtypes = [C.c_char_p]
    assert not lib.tokenizer_create(str(tmp_path / "missing").encode())
    bad = tmp_path / "bad"
    with open(str(bad) + ".tokenizer", "wb") as f:
        f.write(struct.pack("<Iiii", 0x12345678, 2, 4, 4))
    assert not lib.tokenizer_create(str(bad).encode())
    trunc = tmp_path / "trunc"
    with open(str(trunc) + ".tokenizer", "wb") as f:
        f.write(struct.pack("<Iiii", 0x71746B6E, 2, 4, 4) + struct.pack("<10i", *([-1] * 10)) + struct.pack("<fi", 0.0, 3) + b"ab")
    assert not lib.tokenizer_create(str(trunc).encode())
