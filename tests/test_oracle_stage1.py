"""Pins oracle/vf_oracle.c (stage 1) against fixtures produced by the REFERENCE's
own code (tests/golden/make_stage1_golden.py) and the KATs of SURVEY.md §8c."""
import os

import numpy as np
import pytest

from oracle import stage1

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage1_golden.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def bpe():
    return stage1.OracleBPE()


def test_kats(bpe):
    # SURVEY.md §8(c) derived KATs (reference BPEEncoder + tokenizers 0.22.2)
    assert bpe.encode("ACGTRYNNacgtAAAAAGGGCTTCAGxxCTGTGG").tolist() == [80, 14, 12, 17, 80, 146, 60, 499, 498]
    assert bpe.encode("AAAAA").tolist() == [45, 4]
    assert bpe.encode("AAAA").tolist() == [45]
    assert bpe.encode("TTTTTTT").tolist() == [291, 14]
    assert bpe.encode("").tolist() == [] and bpe.encode("NNNN").tolist() == []


def test_bpe_matches_reference_fuzz(bpe):
    seqs, flat, off = G["seqs"], G["tok_flat"], G["tok_off"]
    for i, s in enumerate(seqs):
        want = flat[off[i]:off[i + 1]].astype(np.int32)
        got = bpe.encode(str(s))
        assert got.shape == want.shape and (got == want).all(), f"sequence {i} (len {len(s)})"


def test_token_at_matches_reference(bpe):
    for si, p, want in zip(G["pos_seq_idx"], G["pos_pos"], G["pos_tok"]):
        assert bpe.token_at(str(G["seqs"][si]), int(p)) == int(want)


def test_reverse_complement_matches_reference():
    for s, want in zip(G["seqs"], G["rc"]):
        assert stage1.reverse_complement(str(s)) == str(want)


def test_iupac_table_matches_reference():
    bases = str(G["het_bases"])
    for i, a in enumerate(bases):
        for j, b in enumerate(bases):
            assert stage1.iupac_het(a, b) == chr(G["het_tbl"][i, j])


def test_single_base_variant_matches_reference_apply_variant():
    # vepdataset.SequenceProcessor.apply_variant: het -> IUPAC code, hom -> ALT spliced over ONE base
    for s, p, ref, alt, het, hom in zip(G["av_in"], G["av_pos"], G["av_ref"], G["av_alt"], G["av_het"], G["av_hom"]):
        s, alt = str(s), str(alt)
        chrom = np.frombuffer(s.encode(), np.uint8)
        pool = np.frombuffer(alt.encode(), np.uint8)
        hom_fwd = stage1.apply_variants(chrom, 0, len(s), [p], [1], [0], [len(alt)], [2], pool).decode()
        assert hom_fwd == str(hom).split(",")[0]
        assert stage1.reverse_complement(hom_fwd) == str(hom).split(",")[1]
        if len(alt) == 1:
            het_fwd = stage1.apply_variants(chrom, 0, len(s), [p], [1], [0], [1], [1], pool).decode()
            assert het_fwd == str(het).split(",")[0]


def test_adjust_and_chunkify():
    ids = np.arange(1, 451, dtype=np.int32)
    o, m = stage1.adjust_length(ids[:37])
    assert o[:37].tolist() == ids[:37].tolist() and (o[37:] == 0).all() and (~m[:37]).all() and m[37:].all()
    o, m = stage1.adjust_length(ids)
    assert o.tolist() == ids[:200].tolist() and not m.any()
    c, cm = stage1.chunkify(ids)
    assert c.shape == (3, 200) and c[2, :50].tolist() == ids[400:].tolist() and (c[2, 50:] == 0).all()
    assert cm[2, 50:].all() and not cm[:2].any()
    c, cm = stage1.chunkify(ids, max_chunks=2)
    assert c.shape == (2, 200)
    c, cm = stage1.chunkify(ids[:400])
    assert c.shape == (2, 200) and not cm.any()


def test_windows():
    # utils/data_process.py:21-24 and :387-400
    assert stage1.cre_window(30, 380, 50) == (0, 430)
    assert stage1.cre_window(1000, 1350, 50) == (950, 1400)
    assert stage1.gene_window(5000, 9000, False) == (4000, 9000)
    assert stage1.gene_window(5000, 905000, False) == (4000, 304000)       # uses the shifted start
    assert stage1.gene_window(5000, 9000, True) == (5000, 10000)
    assert stage1.gene_window(5000, 905000, True) == (605000, 906000)


def test_indel_policy_and_overlap():
    ref = np.frombuffer(b"ACGTACGTACGT", np.uint8)
    pool = np.frombuffer(b"TTTG", np.uint8)
    # deletion of "GTA" -> "G" (pos 2, ref_len 3, alt "G"), then a SNP overlapping the deletion is skipped
    out = stage1.apply_variants(ref, 0, 12, [2, 3, 8], [3, 1, 1], [3, 0, 0], [1, 1, 1], [2, 2, 1], pool)
    assert out == b"ACGCGTWCGT"
    # insertion A -> ATTT at pos 4 (het: ALT applied, documented policy)
    out = stage1.apply_variants(ref, 0, 12, [4], [1], [0], [3], [1], np.frombuffer(b"ATT", np.uint8))
    assert out == b"ACGTATTCGTACGT"
    # snp_only drops the indel
    out = stage1.apply_variants(ref, 0, 12, [4], [1], [0], [3], [1], np.frombuffer(b"ATT", np.uint8), snp_only=True)
    assert out == b"ACGTACGTACGT"
    # window clipping: variant outside / straddling the window is ignored; lower-case REF het upper-cases for the code
    ref2 = np.frombuffer(b"acgtacgtacgt", np.uint8)
    out = stage1.apply_variants(ref2, 4, 8, [1, 5, 7], [1, 1, 3], [0, 1, 2], [1, 1, 1], [2, 1, 2], np.frombuffer(b"TTG", np.uint8))
    assert out == b"aYgt"
