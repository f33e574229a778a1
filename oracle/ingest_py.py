"""TEST INFRASTRUCTURE — pure-Python restatement of the host ingest (FASTA -> byte arrays, one sample of a VCF ->
sorted variant arrays), the checker for libvf_ingest.so (variantformer_b200/csrc/vf_ingest.cpp, include/vf_ingest.h).

It states the semantics of what the reference reads through `samtools faidx` / `bcftools consensus` per window
(utils/data_process.py:27,40-59): symbolic ALTs excluded, hom-ref / missing genotypes dropped, het between two
different ALT SNPs resolved to its IUPAC code (datasets/vepdataset.py:75-92).  Only tests/ may import it.
"""
import gzip

import numpy as np

_IUPAC2 = {frozenset("AC"): "M", frozenset("AG"): "R", frozenset("AT"): "W", frozenset("CG"): "S",
           frozenset("CT"): "Y", frozenset("GT"): "K"}


def _open(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def load_fasta(path, chroms=None):
    """-> {name: uint8 array} with the FASTA's own case (soft-masking) preserved."""
    out, name, parts = {}, None, []

    def flush():
        if name is not None and (chroms is None or name in chroms):
            out[name] = np.frombuffer(b"".join(parts), np.uint8).copy()
    with _open(path) as f:
        for line in f:
            if line.startswith(b">"):
                flush()
                name = line[1:].split()[0].decode(); parts = []
            elif chroms is None or name in chroms:
                parts.append(line.rstrip(b"\r\n"))
    flush()
    return out


def load_vcf_sample(path, sample=None, chroms=None):
    """One sample's genotypes -> {chrom: dict(pos int64 0-based, ref_len, alt [bytes], gt uint8)}.
    Records with symbolic ALT (<...>) or '*' are dropped (the reference's `-e 'ALT~"<.*>"'`); hom-ref and missing
    genotypes are dropped; a het between two different ALT SNPs (1/2) is resolved here to its IUPAC code."""
    per = {}
    col = 9
    with _open(path) as f:
        for raw in f:
            if raw.startswith(b"##"):
                continue
            fields = raw.rstrip(b"\r\n").split(b"\t")
            if raw.startswith(b"#CHROM"):
                names = [x.decode() for x in fields[9:]]
                if sample is not None:
                    col = 9 + names.index(sample)
                continue
            chrom = fields[0].decode()
            if chroms is not None and chrom not in chroms:
                continue
            ref = fields[3]; alts = fields[4].split(b",")
            gt_field = fields[col].split(b":")[0] if len(fields) > col else b"./."
            als = gt_field.replace(b"|", b"/").split(b"/")
            if any(a in (b".", b"") for a in als):
                continue
            als = [int(a) for a in als]
            if len(als) == 1:
                als = als * 2
            nz = [a for a in als if a > 0]
            if not nz or max(nz) > len(alts):             # hom-ref, or a malformed call beyond the ALT list: dropped
                continue
            a0 = alts[nz[0] - 1]
            if a0.startswith(b"<") or a0 == b"*":
                continue
            d = per.setdefault(chrom, dict(pos=[], ref_len=[], alt=[], gt=[]))
            if len(set(als)) == 1:
                gt, alt = 2, a0
            elif len(nz) == 2 and len(ref) == 1 and all(len(alts[a - 1]) == 1 for a in nz):
                code = _IUPAC2.get(frozenset((alts[nz[0] - 1] + alts[nz[1] - 1]).decode().upper()))
                gt, alt = 2, (code or "N").encode()
            else:
                gt, alt = 1, a0
            d["pos"].append(int(fields[1]) - 1); d["ref_len"].append(len(ref)); d["alt"].append(alt); d["gt"].append(gt)
    for chrom, d in per.items():
        order = np.argsort(np.asarray(d["pos"], np.int64), kind="stable")
        d["pos"] = np.asarray(d["pos"], np.int64)[order]
        d["ref_len"] = np.asarray(d["ref_len"], np.int32)[order]
        d["gt"] = np.asarray(d["gt"], np.uint8)[order]
        d["alt"] = [d["alt"][i] for i in order]
    return per
