"""Generate tests/golden/stage1_golden.npz from the REFERENCE's own code.

Run in the build container only (needs /root/reference and the `tokenizers`
wheel):  python tests/golden/make_stage1_golden.py

What is recorded (all produced by unmodified reference functions):
  * utils/seq.py::BPEEncoder.encode        -> token ids for fuzzed sequences
  * utils/seq.py::BPEEncoder.encode_with_position -> covering-token index
  * utils/functions.py::reverse_complement -> strings
  * datasets/vepdataset.py::SequenceProcessor.get_iupac_code / apply_variant
The fixture is what pins oracle/vf_oracle.c (tests/test_oracle_stage1.py) and,
through it, the CUDA stage-1 kernels on the GPU box.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402


def main():
    refshim.install()
    for name in ("duckdb", "fsspec"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    from utils.seq import BPEEncoder
    from utils.functions import reverse_complement
    from datasets.vepdataset import SequenceProcessor, Variant

    enc = BPEEncoder()
    enc.load_vocabulary(os.path.join(refshim.REF_ROOT, "vocabs", "bpe_vocabulary_500.json"))
    rng = np.random.default_rng(20240601)

    def rand_seq(n, kind):
        if kind == 0:      # plain ACGT
            s = rng.choice(list("ACGT"), n)
        elif kind == 1:    # full IUPAC + N + soft-mask + junk
            s = rng.choice(list("ACGTACGTACGTRYSWKMBDHVNacgtn-."), n)
        elif kind == 2:    # homopolymer rich
            s = rng.choice(list("AATTTTAAAACG"), n)
            runs = rng.integers(0, max(n - 12, 1), 4)
            for r in runs:
                s[r:r + rng.integers(2, 12)] = rng.choice(list("AT"))
        else:              # dinucleotide repeats + N runs
            unit = "".join(rng.choice(list("ACGT"), rng.integers(1, 5)))
            s = np.array(list((unit * (n // len(unit) + 1))[:n]))
            if n > 20:
                p = rng.integers(0, n - 10); s[p:p + rng.integers(1, 10)] = "N"
        return "".join(s)

    seqs = ["ACGTRYNNacgtAAAAAGGGCTTCAGxxCTGTGG", "AAAAA", "AAAA", "TTTTTTT", "A", "N", "", "NNNN", "acgtn"]
    for kind in range(4):
        for n in list(rng.integers(1, 500, 60)) + [1000, 2500]:
            seqs.append(rand_seq(int(n), kind))
    seqs.append(rand_seq(60000, 0))           # one long word (gene-window scale, shortened)
    seqs.append(rand_seq(30000, 2))

    tok_flat, tok_off = [], [0]
    for s in seqs:
        ids, _, _, _ = enc.encode([s, "A"])
        tok_flat.extend(ids); tok_off.append(len(tok_flat))

    # encode_with_position on a subset
    pos_seq_idx, pos_pos, pos_tok = [], [], []
    for si, s in enumerate(seqs[:200]):
        if len(s) == 0:
            continue
        for p in rng.integers(0, len(s), 3):
            try:
                r = enc.encode_with_position(s, int(p))["position_id"]
            except ValueError:
                r = -2
            pos_seq_idx.append(si); pos_pos.append(int(p)); pos_tok.append(r)

    rc = [reverse_complement(s) for s in seqs[:120]]

    bases = "ACGTNRacgt"
    het_tbl = np.array([[ord(SequenceProcessor.get_iupac_code(a, b)) for b in bases] for a in bases], np.uint8)

    # apply_variant (single-base replacement semantics of the VEP path)
    av_in, av_pos, av_ref, av_alt, av_het, av_hom = [], [], [], [], [], []
    for _ in range(40):
        s = rand_seq(int(rng.integers(20, 120)), 0)
        p = int(rng.integers(0, len(s)))
        ref = s[p]; alt = str(rng.choice([c for c in "ACGT" if c != ref]))
        if rng.random() < 0.25:
            alt = alt + "".join(rng.choice(list("ACGT"), rng.integers(1, 4)))   # insertion-like ALT
        v = Variant(chrom="chr1", pos=p + 1, ref=ref, alt=alt, tissue="x", gene_id=["g"])
        het, hom = SequenceProcessor.apply_variant(s + "," + s, v, p)
        av_in.append(s); av_pos.append(p); av_ref.append(ref); av_alt.append(alt)
        av_het.append(het); av_hom.append(hom)

    out = os.path.join(HERE, "stage1_golden.npz")
    np.savez_compressed(
        out,
        seqs=np.array(seqs, dtype=object), tok_flat=np.asarray(tok_flat, np.int16), tok_off=np.asarray(tok_off, np.int64),
        pos_seq_idx=np.asarray(pos_seq_idx, np.int32), pos_pos=np.asarray(pos_pos, np.int32), pos_tok=np.asarray(pos_tok, np.int32),
        rc=np.array(rc, dtype=object), het_bases=np.array(bases), het_tbl=het_tbl,
        av_in=np.array(av_in, dtype=object), av_pos=np.asarray(av_pos, np.int32), av_ref=np.array(av_ref, dtype=object),
        av_alt=np.array(av_alt, dtype=object), av_het=np.array(av_het, dtype=object), av_hom=np.array(av_hom, dtype=object),
    )
    print("wrote", out, os.path.getsize(out), "bytes;", len(seqs), "sequences,", len(tok_flat), "tokens")


if __name__ == "__main__":
    main()
