"""CPU-side checks: the C-ABI library loads and exports every symbol include/vf_b200.h declares; product code has no
CPU fallback; ingest parsers; config surface."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from variantformer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vf_b200.h")).read()
    declared = set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                                    # dlopen only, no device access
    for name in declared:
        assert hasattr(lib, name)
    assert lib.vf_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    from variantformer_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less machine")
    with pytest.raises(_lib.VFError, match="no CPU fallback"):
        _lib.lib()
    from variantformer_b200.processors.model_manager import ModelManager
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ModelManager({"model_class": "Seq2GenePredictorCombinedModulator"})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "variantformer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "fake_ops" not in src


def test_ingest_fasta_and_vcf(tmp_path):
    from variantformer_b200 import ingest
    fa = tmp_path / "g.fa"
    fa.write_text(">chr1 desc\nACGTacgtNN\nGGCC\n>chr2\nTTTT\n")
    g = ingest.load_fasta(str(fa))
    assert g["chr1"].tobytes() == b"ACGTacgtNNGGCC" and g["chr2"].tobytes() == b"TTTT"
    assert list(ingest.load_fasta(str(fa), chroms={"chr2"})) == ["chr2"]
    vcf = tmp_path / "s.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
                   "chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1\t1/1\n"
                   "chr1\t2\t.\tC\tT,G\t.\t.\t.\tGT:DP\t1|2:9\t0/0\n"
                   "chr1\t7\t.\tGT\tG\t.\t.\t.\tGT\t1/1\t./.\n"
                   "chr1\t9\t.\tN\t<DEL>\t.\t.\t.\tGT\t1/1\t1/1\n"
                   "chr1\t11\t.\tG\tGAA\t.\t.\t.\tGT\t0|1\t0/0\n"
                   "chr2\t1\t.\tT\tC\t.\t.\t.\tGT\t0/0\t0/1\n")
    v = ingest.load_vcf_sample(str(vcf), sample="S1")
    c1 = v["chr1"]
    assert c1["pos"].tolist() == [1, 4, 6, 10] and c1["ref_len"].tolist() == [1, 1, 2, 1]
    assert c1["alt"] == [b"K", b"G", b"G", b"GAA"] and c1["gt"].tolist() == [2, 1, 2, 1]     # 1|2 het SNP -> IUPAC K
    assert "chr2" not in v
    v2 = ingest.load_vcf_sample(str(vcf), sample="S2")
    assert v2["chr1"]["pos"].tolist() == [4] and v2["chr1"]["gt"].tolist() == [2] and v2["chr2"]["gt"].tolist() == [1]


def test_config_surface():
    from variantformer_b200.utils.config import CONFIG_DIR, Config, load_yaml
    cfg = load_yaml(os.path.join(CONFIG_DIR, "vf_model.yaml"))
    m = cfg.v4_pcg.model
    assert m.model_class == "Seq2GenePredictorCombinedModulator" and m.emb_dim == 1536 and m["num_layers"] == 25
    c = m.copy(); del c["cre_tokenizer"]; delattr(c, "gene_tokenizer")
    assert "cre_tokenizer" in m and "cre_tokenizer" not in c and c.get("nope", 3) == 3
    assert cfg.v4_pcg.dataset.max_chunks == 200 and cfg.v4_ag.model.checkpoint_path.endswith("v4_ag_epoch9_checkpoint.pth")
    assert isinstance(Config({"a": {"b": 1}}).a, Config)


def test_merge_table_property_and_hf_loader():
    from variantformer_b200.stage1 import load_merge_table
    a, b, c, vocab = load_merge_table()
    assert len(a) == 482 and len(vocab) == 500 and (c == np.arange(18, 500)).all()
    assert (a < c).all() and (b < c).all()               # operands pre-exist: the rank-sweep kernel's precondition
    assert int((a == b).sum()) == 27                     # self-pair merges


def test_window_arithmetic_matches_oracle():
    from oracle import stage1 as O
    from variantformer_b200.stage1 import cre_window, gene_window
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = int(rng.integers(0, 2_000_000)); e = s + int(rng.integers(1, 900_000))
        assert cre_window(s, e, 50) == O.cre_window(s, e, 50)
        for strand in "+-":
            assert gene_window(s, e, strand, 1000, 300000) == O.gene_window(s, e, strand == "-")


def test_attention_work_items_cover_every_row_once():
    """Work tables of vf_attention_mc_varlen (include/vf_b200.h).  Default: left-over tiles are paired even when they
    read different keys (split items; the kernel orders its waits behind the producer's issue counters).  With pairing
    off (VF_PAIR_TILES=0) the two slots of an item read the same keys or the second slot is empty.  Either way every
    query row appears exactly once and slot 0 of an item is never the empty one."""
    from variantformer_b200 import ops
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 700, 300).tolist() + [97] * 40 + [128, 129, 256, 257]
    units = np.array([[0, 128, 0, 200, 0], [128, 72, 0, 200, 128], [200, 1, 500, 201, 3], [201, 1, 701, 201, 9],
                      [202, 63, 1000, 1024, 0], [265, 63, 1000, 1024, 0], [328, 63, 1000, 1024, 0]])
    saved = ops.PAIR_UNRELATED_TILES
    try:
        for pairing in (True, False):
            ops.PAIR_UNRELATED_TILES = pairing
            sm = ops.SlotMap(lens, "cpu")
            tab = sm.table.numpy()
            both = (tab[:, 0, 1] > 0) & (tab[:, 1, 1] > 0)
            assert both.any() and np.all(tab[:, 0, 1] > 0)
            split = both & ((tab[:, 0, 2] != tab[:, 1, 2]) | (tab[:, 0, 3] != tab[:, 1, 3]))
            assert split.any() == pairing
            if not pairing:
                assert (~both).any() and np.all(tab[~both, 1, :] == 0)
            else:
                assert (~both).sum() <= 1                              # at most one tile is left without a partner
            rows = np.concatenate([np.arange(r[0], r[0] + r[1]) for r in tab.reshape(-1, 8) if r[1] > 0])
            assert len(rows) == sum(lens) and len(np.unique(rows)) == len(rows)
            fu = ops.SlotMap.from_units(units, "cpu").table.numpy()
            both = (fu[:, 0, 1] > 0) & (fu[:, 1, 1] > 0)
            if pairing:
                assert fu.shape[0] == 4 and both.tolist() == [True, True, True, False]
            else:                                                      # consecutive units pair only on the same key range
                assert fu.shape[0] == 5 and both.tolist() == [True, False, False, True, False]
                assert np.all(fu[both, 0, 2:4] == fu[both, 1, 2:4])
            got = np.concatenate([fu[:, 0, :5][fu[:, 0, 1] > 0], fu[:, 1, :5][fu[:, 1, 1] > 0]])
            assert sorted(map(tuple, got)) == sorted(map(tuple, units))
    finally:
        ops.PAIR_UNRELATED_TILES = saved
    # sequences without keys: no work item, their output rows are listed for zero-filling
    sm = ops.SlotMap([5, 130, 7], "cpu", k_lens=[9, 0, 4])
    assert sm.keyless == [(5, 135)] and sm.n_items == 1


def test_attention_build_slots_equals_the_python_tables():
    from variantformer_b200 import _lib, ops
    lib = _lib.load()
    rng = np.random.default_rng(2)
    for q_lens, k_lens in [(rng.integers(1, 700, 200), None), ([97] * 33, None), ([12663] * 3, [1024, 5, 300]),
                           ([5, 130, 7, 0, 64], [9, 0, 4, 3, 1])]:
        q = np.asarray(q_lens, np.int32); k = None if k_lens is None else np.asarray(k_lens, np.int32)
        for pairing in (1, 0):
            ops.PAIR_UNRELATED_TILES, saved = bool(pairing), ops.PAIR_UNRELATED_TILES
            try:
                want = ops.SlotMap(q, "cpu", k_lens=k).table.numpy()
            finally:
                ops.PAIR_UNRELATED_TILES = saved
            n = lib.vf_attention_build_slots(q.ctypes.data, None if k is None else k.ctypes.data, len(q), pairing, None, 0)
            assert n == want.shape[0]
            got = np.zeros((n, 2, 8), np.int32)
            assert lib.vf_attention_build_slots(q.ctypes.data, None if k is None else k.ctypes.data, len(q), pairing,
                                                got.ctypes.data, n) == n
            assert np.array_equal(got, want)
