"""CPU ORACLE (test infrastructure, NOT product code) — stage 1 wrappers.

numpy/ctypes front-end of oracle/vf_oracle.c: genotype application, reverse
complement, BPE-500 tokenisation, pad/truncate and chunking.  Mirrors the call
sequence of the reference's datasets/vcfdataset.py:219-303 so parity tests read
like the reference's own data path.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs import this module.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_MERGES = os.path.join(_HERE, "..", "variantformer_b200", "vocabs", "bpe500_merges.txt")
ALPHABET = "ABCDGHKMRSTVWY"  # ids 4..17 (vocabs/bpe_vocabulary_500.json)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.vfo_bpe_create.restype = vp
        L.vfo_bpe_create.argtypes = [vp, vp, vp, i32, i32]
        L.vfo_bpe_destroy.argtypes = [vp]
        L.vfo_bpe_encode.restype = i32
        L.vfo_bpe_encode.argtypes = [vp, vp, i32, vp, vp]
        L.vfo_bpe_token_at.restype = i32
        L.vfo_bpe_token_at.argtypes = [vp, vp, i32, i32]
        L.vfo_reverse_complement.argtypes = [vp, i32, vp]
        L.vfo_iupac_het.restype = C.c_uint8
        L.vfo_iupac_het.argtypes = [C.c_uint8, C.c_uint8]
        L.vfo_apply_variants.restype = i32
        L.vfo_apply_variants.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, i32, vp, i32, vp]
        L.vfo_adjust_length.argtypes = [vp, i32, i32, vp, vp]
        L.vfo_chunkify.restype = i32
        L.vfo_chunkify.argtypes = [vp, i32, i32, i32, vp, vp]
        L.vfo_cre_window.argtypes = [i64, i64, i64, vp, vp]
        L.vfo_gene_window.argtypes = [i64, i64, i32, i64, i64, vp, vp]
        _lib = L
    return _lib


def load_merges(path=_MERGES):
    """-> (left ids, right ids, new ids) int16 arrays from the rank-ordered merge table."""
    vocab = {"<pad>": 0, "<s>": 1, "</s>": 2, "<unk>": 3}
    for i, ch in enumerate(ALPHABET):
        vocab[ch] = 4 + i
    left, right, new = [], [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            a, b = line.split()
            vocab[a + b] = len(vocab)
            left.append(vocab[a]); right.append(vocab[b]); new.append(vocab[a + b])
    return (np.asarray(left, np.int16), np.asarray(right, np.int16), np.asarray(new, np.int16), vocab)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleBPE:
    """Restates utils/seq.py::BPEEncoder (forward strand only) on top of vf_oracle.c."""

    def __init__(self, merges_path=_MERGES):
        l, r, n, self.vocab = load_merges(merges_path)
        self._h = lib().vfo_bpe_create(_p(l), _p(r), _p(n), len(l), len(self.vocab))
        assert self._h

    def encode(self, seq) -> np.ndarray:
        b = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), np.uint8)
        out = np.empty(max(len(b), 1), np.int32)
        n = lib().vfo_bpe_encode(self._h, _p(b), len(b), _p(out), None)
        assert n >= 0
        return out[:n].copy()

    def encode_with_starts(self, seq):
        b = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), np.uint8)
        out = np.empty(max(len(b), 1), np.int32)
        st = np.empty(max(len(b), 1), np.int32)
        n = lib().vfo_bpe_encode(self._h, _p(b), len(b), _p(out), _p(st))
        return out[:n].copy(), st[:n].copy()

    def token_at(self, seq, position: int) -> int:
        b = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), np.uint8)
        return int(lib().vfo_bpe_token_at(self._h, _p(b), len(b), position))


def reverse_complement(seq):
    b = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), np.uint8)
    out = np.empty_like(b)
    lib().vfo_reverse_complement(_p(b), len(b), _p(out))
    return out.tobytes().decode() if isinstance(seq, str) else out.tobytes()


def iupac_het(ref: str, alt: str) -> str:
    return chr(lib().vfo_iupac_het(ord(ref), ord(alt)))


def apply_variants(chrom_seq: np.ndarray, w0: int, w1: int, pos, ref_len, alt_off, alt_len, gt,
                   alt_pool: np.ndarray, snp_only=False) -> bytes:
    pos = np.ascontiguousarray(pos, np.int64); ref_len = np.ascontiguousarray(ref_len, np.int32)
    alt_off = np.ascontiguousarray(alt_off, np.int32); alt_len = np.ascontiguousarray(alt_len, np.int32)
    gt = np.ascontiguousarray(gt, np.uint8); alt_pool = np.ascontiguousarray(alt_pool, np.uint8)
    out = np.empty(int(w1 - w0) + int(alt_len.sum()) + 1, np.uint8)
    n = lib().vfo_apply_variants(_p(chrom_seq), w0, w1, _p(pos), _p(ref_len), _p(alt_off), _p(alt_len),
                                 _p(gt), len(pos), _p(alt_pool), int(snp_only), _p(out))
    return out[:n].tobytes()


def adjust_length(ids: np.ndarray, max_length=200):
    ids = np.ascontiguousarray(ids, np.int32)
    o = np.empty(max_length, np.int32); m = np.empty(max_length, np.uint8)
    lib().vfo_adjust_length(_p(ids), len(ids), max_length, _p(o), _p(m))
    return o, m.astype(bool)


def chunkify(ids: np.ndarray, max_length=200, max_chunks=200):
    ids = np.ascontiguousarray(ids, np.int32)
    cap = min(max_chunks, (len(ids) + max_length - 1) // max_length)
    o = np.empty((max(cap, 1), max_length), np.int32); m = np.empty((max(cap, 1), max_length), np.uint8)
    g = lib().vfo_chunkify(_p(ids), len(ids), max_length, max_chunks, _p(o), _p(m))
    return o[:g], m[:g].astype(bool)


def cre_window(start, end, nb=50):
    a, b = C.c_int64(), C.c_int64()
    lib().vfo_cre_window(start, end, nb, C.byref(a), C.byref(b))
    return a.value, b.value


def gene_window(start, end, minus, up=1000, down=300000):
    a, b = C.c_int64(), C.c_int64()
    lib().vfo_gene_window(start, end, int(minus), up, down, C.byref(a), C.byref(b))
    return a.value, b.value
