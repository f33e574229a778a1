"""libvf_ingest.so (C++: parallel BGZF inflation + FASTA / VCF parsing) against the pure-Python restatement
oracle/ingest_py.py on seeded random files, in plain, gzip and BGZF form.  CPU only."""
import gzip
import os
import re
import struct
import zlib

import numpy as np
import pytest

from oracle import ingest_py
from variantformer_b200 import ingest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bgzf(data: bytes, block=30000) -> bytes:
    """Write `data` the way bgzip does: gzip members of <= 64 KB with the 'BC' extra field + the empty EOF block."""
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
    for ch in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(ch) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6)
        out += b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(ch) & 0xffffffff, len(ch))
    return bytes(out)


def _write(path, data, form):
    if form == "plain":
        path.write_bytes(data)
    elif form == "gzip":
        path.write_bytes(gzip.compress(data[:len(data) // 2]) + gzip.compress(data[len(data) // 2:]))   # two members
    else:
        path.write_bytes(_bgzf(data))
    return str(path)


def _random_fasta(rng, n_seq=5):
    parts = []
    for i in range(n_seq):
        n = int(rng.integers(1, 200_000))
        seq = rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), n).tobytes()
        width = int(rng.choice([60, 70, 80]))
        eol = b"\r\n" if i == 2 else b"\n"
        parts.append(b">chr%d some description %d" % (i + 1, i) + eol)
        parts += [seq[j:j + width] + eol for j in range(0, n, width)]
    return b"".join(parts)


def _random_vcf(rng, n=20000):
    lines = [b"##fileformat=VCFv4.2", b"##contig=<ID=chr1>",
             b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3"]
    bases = [b"A", b"C", b"G", b"T"]
    gts = [b"0/0", b"0/1", b"1/1", b"1|0", b"0|1", b"1|2", b"2/1", b"./.", b".", b"1", b"0", b"2|2", b"1/1/1", b"0/1/2"]
    for _ in range(n):
        chrom = b"chr%d" % int(rng.integers(1, 4))
        pos = int(rng.integers(1, 5_000_000))
        kind = rng.random()
        if kind < 0.7:
            ref = bases[int(rng.integers(4))]; alt = b",".join(bases[int(x)] for x in rng.choice(4, int(rng.integers(1, 3)), replace=False))
        elif kind < 0.85:
            ref = b"".join(bases[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6)))); alt = bases[int(rng.integers(4))] + b",<DEL>"
        elif kind < 0.95:
            ref = bases[int(rng.integers(4))]; alt = ref + b"".join(bases[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 8)))) + b",*"
        else:
            ref = b"N"; alt = b"<INS>,A"
        cols = [gts[int(rng.integers(len(gts)))] + (b":12:0.5" if rng.random() < 0.5 else b"") for _ in range(3)]
        lines.append(b"\t".join([chrom, str(pos).encode(), b".", ref, alt, b".", b"PASS", b".", b"GT:DP:AF"] + cols))
    return b"\n".join(lines) + b"\n"


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vf_ingest.h")).read()
    declared = set(re.findall(r"\b(vf_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(ingest.SIGNATURES), declared ^ set(ingest.SIGNATURES)
    lib = ingest.lib()
    for name in declared:
        assert hasattr(lib, name)


@pytest.mark.parametrize("form", ["plain", "gzip", "bgzf"])
def test_fasta_matches_python_restatement(tmp_path, form):
    rng = np.random.default_rng(11)
    path = _write(tmp_path / f"g.{form}", _random_fasta(rng), form)
    want = ingest_py.load_fasta(path)
    got = ingest.load_fasta(path, threads=4)
    assert list(got) == list(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    assert list(ingest.load_fasta(path, chroms={"chr3"})) == ["chr3"]


@pytest.mark.parametrize("form", ["plain", "gzip", "bgzf"])
@pytest.mark.parametrize("sample", [None, "S1", "S3"])
def test_vcf_matches_python_restatement(tmp_path, form, sample):
    rng = np.random.default_rng(5)
    path = _write(tmp_path / f"s.{form}", _random_vcf(rng), form)
    want = ingest_py.load_vcf_sample(path, sample=sample)
    got = ingest.load_vcf_sample(path, sample=sample, threads=3)
    assert sorted(got) == sorted(want)
    for c in want:
        for k in ("pos", "ref_len", "gt"):
            assert np.array_equal(got[c][k], want[c][k]), (c, k)
        assert got[c]["alt"] == want[c]["alt"], c


def test_errors_are_reported(tmp_path):
    with pytest.raises(ingest.IngestError, match="cannot open"):
        ingest.load_fasta(str(tmp_path / "missing.fa"))
    bad = tmp_path / "bad.gz"
    bad.write_bytes(b"\x1f\x8b\x08\x00" + b"\x00" * 30)
    with pytest.raises(ingest.IngestError):
        ingest.load_vcf_sample(str(bad))
    vcf = tmp_path / "s.vcf"
    vcf.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\nchr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1\n")
    with pytest.raises(ingest.IngestError, match="not found"):
        ingest.load_vcf_sample(str(vcf), sample="nobody")
