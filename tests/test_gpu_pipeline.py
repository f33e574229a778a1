"""Full four-stage path on the GPU box: window tables -> tokens (bit-exact) -> expression/embeddings (bf16 tolerance),
through HotPath and through the reference-shaped VCFProcessor / VCFDataset / ModelManager / Trainer surface."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model_fp32, stage1 as O  # noqa: E402  (checker only)
from tests.common import pearson, rel_err  # noqa: E402
from variantformer_b200.engine import Engine  # noqa: E402
from variantformer_b200.pipeline import GeneSpec, HotPath  # noqa: E402
from variantformer_b200.stage1 import Genome, SampleVariants, cre_window, gene_window  # noqa: E402
from variantformer_b200.utils import random_init, synth  # noqa: E402
from variantformer_b200.utils.constants import REF_CREs  # noqa: E402

CFG = dict(random_init.V4_PCG_MODEL, emb_dim=384, gene_emb_dim=256, num_heads=8, num_layers=3, token_dim=256)
HP = dict(random_init.SEQ2REG_HP, embedding_dim=256, num_heads=4, num_layers=2)


def _world(seed=77, chrom_len=3_000_000, n_genes=3, n_cres=40):
    rng = np.random.default_rng(seed)
    chrom = synth.make_chromosome(rng, chrom_len)
    var = synth.make_variants(rng, chrom)
    genes = []
    for g in range(n_genes):
        lay = synth.make_gene_layout(rng, chrom_len, n_cres, strand="+-"[g % 2], body_len=int(rng.integers(20_000, 60_000)),
                                     cre_span=200_000)
        genes.append(GeneSpec("chr1", lay["start"], lay["end"], lay["strand"], lay["cre_start"], lay["cre_end"],
                              lay["labels"], [62, 3, 14][: 1 + g]))
    return {"chr1": chrom}, {"chr1": var}, genes


def _oracle_batch(chroms, var, genes, use_var=True):
    bpe = O.OracleBPE()
    b = {k: [] for k in ("cre_sequences", "cre_attention_masks", "gene_embeddings", "gene_attention_masks",
                         "tissue_context", "ref_cre_labels")}
    for g in genes:
        chrom = chroms[g.chrom]; v = var[g.chrom]
        lens = np.array([len(a) for a in v["alt"]], np.int32); pool = np.frombuffer(b"".join(v["alt"]), np.uint8)
        aoff = np.cumsum(lens) - lens
        minus = g.strand == "-"

        def window(a0, a1):
            if use_var:
                lo, hi = np.searchsorted(v["pos"], a0), np.searchsorted(v["pos"], a1)
                s = O.apply_variants(chrom, a0, a1, v["pos"][lo:hi], v["ref_len"][lo:hi], aoff[lo:hi], lens[lo:hi],
                                     v["gt"][lo:hi], pool)
            else:
                s = chrom[a0:a1].tobytes()
            return O.reverse_complement(s) if minus else s
        order = np.argsort(g.cre_start, kind="stable")
        order = order[::-1] if minus else order
        toks, masks = zip(*[O.adjust_length(bpe.encode(window(*cre_window(g.cre_start[i], g.cre_end[i], 50))), 200)
                            for i in order])
        gt, gm = O.chunkify(bpe.encode(window(*gene_window(g.start, g.end, g.strand, 1000, 300000))), 200, 200)
        b["cre_sequences"].append(torch.from_numpy(np.stack(toks)).long().unsqueeze(1))
        b["cre_attention_masks"].append(torch.from_numpy(np.stack(masks)).unsqueeze(1))
        b["gene_embeddings"].append(torch.from_numpy(gt).long().unsqueeze(1))
        b["gene_attention_masks"].append(torch.from_numpy(gm).unsqueeze(1))
        b["tissue_context"].append(torch.tensor(g.tissues)); b["ref_cre_labels"].append(torch.from_numpy(np.asarray(g.cre_labels)[order].copy()))
    return b


def test_hotpath_tokens_bit_exact_and_outputs_within_tolerance():
    chroms, var, genes = _world()
    sd = random_init.make_state_dict(CFG, HP, seed=5)
    hot = HotPath(Engine(sd, CFG, HP), Genome.from_arrays(chroms, "cuda"))
    sv = SampleVariants(var, "cuda")
    t = hot.tokenize(genes, sv)
    want = _oracle_batch(chroms, var, genes)
    for g in range(len(genes)):
        assert torch.equal(t["cre_tok"][g].cpu().long(), want["cre_sequences"][g][:, 0]), f"CRE tokens gene {g}"
        assert torch.equal(t["cre_msk"][g].cpu(), want["cre_attention_masks"][g][:, 0])
        assert torch.equal(t["gene_tok"][g].cpu().long(), want["gene_embeddings"][g][:, 0]), f"gene tokens gene {g}"
        assert torch.equal(t["gene_msk"][g].cpu(), want["gene_attention_masks"][g][:, 0])
    pred, emb = hot.predict(genes, sv)
    ref = model_fp32.predict_step(sd, CFG, HP, want, schedule="reference")
    w_emb = np.concatenate(ref["embeddings"]); w_pred = np.concatenate(ref["pred_gene_exp"]).ravel()
    assert rel_err(emb, w_emb) <= 1e-2 and pearson(emb, w_emb) >= 0.9999 and rel_err(pred, w_pred) <= 1e-2


def test_pipelined_api_and_stream_overlap_are_invisible(monkeypatch):
    """predict_pipelined (stage 1 of slab i+1 on a side stream, CRE stack on its own stream) returns bit-identical
    results to one-slab-at-a-time, single-stream execution."""
    from variantformer_b200 import engine as engine_mod
    chroms, var, genes = _world()
    sd = random_init.make_state_dict(CFG, HP, seed=5)
    hot = HotPath(Engine(sd, CFG, HP), Genome.from_arrays(chroms, "cuda"))
    sv = SampleVariants(var, "cuda")
    slabs = [genes, genes[::-1], genes[:1], genes]
    monkeypatch.setattr(engine_mod, "CRE_STREAM", False)
    want = [hot.predict(g, sv) for g in slabs]
    monkeypatch.setattr(engine_mod, "CRE_STREAM", True)
    got = list(hot.predict_pipelined(iter(slabs), sv, to_host=True))
    assert len(got) == len(want)
    for (p0, e0), (p1, e1) in zip(want, got):
        assert np.array_equal(p0, p1) and np.array_equal(e0, e1)


def _write_artifacts(tmp_path, chroms, var, genes):
    art = tmp_path / "_artifacts"; (art / "gene_cre_manifests").mkdir(parents=True)
    with open(art / "GRCh38_no_alt_analysis_set_GCA_000001405.15.fasta.gz", "wb") as f:     # plain text is accepted too
        f.write(b">chr1 synthetic\n")
        s = chroms["chr1"].tobytes()
        for i in range(0, len(s), 60):
            f.write(s[i:i + 60] + b"\n")
    v = var["chr1"]
    with open(tmp_path / "sample.vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n")
        for p, rl, alt, gt in zip(v["pos"], v["ref_len"], v["alt"], v["gt"]):
            ref = chroms["chr1"][p:p + rl].tobytes().decode().upper()
            f.write(f"chr1\t{p + 1}\t.\t{ref}\t{alt.decode()}\t.\tPASS\t.\tGT\t{'0/1' if gt == 1 else '1|1'}\n")
    rows = []
    for i, g in enumerate(genes):
        gid = f"ENSG{i:011d}.1"
        rows.append(dict(gene_id=gid, gene_name=f"G{i}", chromosome="chr1", start=g.start, end=g.end, strand=g.strand))
        pd.DataFrame(dict(chromosome="chr1", start_cre=g.cre_start, end_cre=g.cre_end,
                          cre_name=[REF_CREs[l] for l in g.cre_labels])).to_csv(art / "gene_cre_manifests" / f"{gid}.csv", index=False)
    pd.DataFrame(rows).to_csv(art / "all_genes_v1_pcg_gencodeV24.csv", index=False)
    return art, rows


def test_predict_to_parquet_is_resumable_and_equals_predict(tmp_path):
    """The production loop: pinned double-buffered D2H + background Parquet writer, one file per slab; an interrupted
    run resumes from the files that exist and ends with the same bits as HotPath.predict."""
    from variantformer_b200.writer import ResultWriter
    chroms, var, genes = _world(seed=81, n_genes=5, n_cres=24)
    sd = random_init.make_state_dict(CFG, HP, seed=9)
    hot = HotPath(Engine(sd, CFG, HP), Genome.from_arrays(chroms, "cuda"))
    variants = SampleVariants(var, "cuda")
    slabs = [genes[:2], genes[2:3], genes[3:]]
    want = [hot.predict(g, variants) for g in slabs]
    out = str(tmp_path / "run")
    w = ResultWriter(out, CFG["emb_dim"], meta={"seed": 9})
    assert hot.predict_to_parquet(slabs, variants, w, sample="S1") == 3
    w.close()
    assert w.done() == {0, 1, 2}
    os.remove(w.path(1))                                          # "crash" after slab 0 and 2 were written
    w = ResultWriter(out, CFG["emb_dim"], meta={"seed": 9})
    assert w.done() == {0, 2} and hot.predict_to_parquet(slabs, variants, w, sample="S1") == 1
    assert hot.predict_to_parquet(slabs, variants, w, sample="S1") == 0
    w.close()
    df = w.read_all()
    pred = np.concatenate([p for p, _ in want]); emb = np.concatenate([e for _, e in want])
    assert len(df) == len(pred) and np.array_equal(df["predicted_expression"].to_numpy(), pred)
    assert np.array_equal(np.stack(df["embeddings"].to_numpy()), emb)
    assert df["tissue"].tolist() == [t for g in genes for t in g.tissues] and set(df["sample"]) == {"S1"}


def test_vcfprocessor_surface_end_to_end(tmp_path):
    from variantformer_b200.processors.vcfprocessor import VCFProcessor
    chroms, var, genes = _world(seed=78, n_genes=2, n_cres=24)
    art, rows = _write_artifacts(tmp_path, chroms, var, genes)
    over = dict(CFG, random_init=5, seq2reg_hyper_parameters=HP)
    vp = VCFProcessor("v4_pcg", base_dir=str(tmp_path), model_overrides=over)
    assert "whole blood" in vp.get_tissues() and len(vp.get_genes()) == 2
    query = pd.DataFrame({"gene_id": [r["gene_id"] for r in rows] + ["ENSG_missing"],
                          "tissues": ["whole blood,K562,not-a-tissue", "thyroid", "whole blood"]})
    ds, dl = vp.create_data(str(tmp_path / "sample.vcf"), query, batch_size=2)
    model, ckpt, trainer = vp.load_model()
    assert model.vep is False and not model.training and trainer.precision == "bf16-mixed"
    out = vp.predict(model, None, trainer, dl, ds)
    assert list(out["gene_id"]) == [r["gene_id"] for r in rows]
    assert out["predicted_expression"].iloc[0].shape == (2, 1) and out["embeddings"].iloc[0].shape == (2, CFG["emb_dim"])
    # oracle on the same genome / variants / weights
    tv = vp.tissue_vocab
    for g, names in zip(genes, (["whole blood", "K562"], ["thyroid"])):
        g.tissues = [tv[n] for n in names]
    sd = random_init.make_state_dict({k: v for k, v in CFG.items()}, HP, seed=5)
    ref = model_fp32.predict_step(sd, CFG, HP, _oracle_batch(chroms, var, genes), schedule="reference")
    for i in range(2):
        assert rel_err(out["embeddings"].iloc[i], ref["embeddings"][i]) <= 1e-2
        assert pearson(out["embeddings"].iloc[i], ref["embeddings"][i]) >= 0.9999
        assert rel_err(out["predicted_expression"].iloc[i], ref["pred_gene_exp"][i]) <= 1e-2
    # checkpoint path: tokenizer checkpoints load ({"hyper_parameters","state_dict"}), missing model checkpoint raises
    tok_sd = {k[len("cre_tokenizer."):]: v for k, v in sd.items() if k.startswith("cre_tokenizer.")}
    torch.save({"hyper_parameters": HP, "state_dict": tok_sd}, art / "pretrained_tokenizers_checkpoint.pth")
    with pytest.raises(ValueError, match="Checkpoint not found"):
        VCFProcessor("v4_pcg", base_dir=str(tmp_path), model_overrides=dict(CFG)).load_model()
    torch.save({"state_dict": sd}, art / "v4_pcg_epoch11_checkpoint.pth")
    vp2 = VCFProcessor("v4_pcg", base_dir=str(tmp_path), model_overrides=dict(CFG))
    model2, ckpt2, trainer2 = vp2.load_model()
    ds2, dl2 = vp2.create_data(str(tmp_path / "sample.vcf"), query, batch_size=1)
    out2 = vp2.predict(model2, ckpt2, trainer2, dl2, ds2)
    for i in range(2):           # same weights through the checkpoint path, different batching -> identical
        assert np.array_equal(out2["embeddings"].iloc[i], out["embeddings"].iloc[i])


def test_vep_triplet_matches_reference_semantics():
    """(ref, het, hom) batches vs the oracle applied to Python strings built with the reference's own rule
    (vepdataset.py:94-131: het = one base -> IUPAC code or N, hom = one base -> ALT string), and the full
    variant_prediction surface of the model."""
    from variantformer_b200.datasets.vepdataset import VEPBatchBuilder, Variant, get_iupac_code
    from variantformer_b200.seq2gene.model_combined_modulator import Seq2GenePredictorCombinedModulator, attach_trainer
    from variantformer_b200.seq2reg.model import Seq2RegPredictor
    chroms, var, genes = _world(seed=79, n_genes=2, n_cres=30)
    genome = Genome.from_arrays(chroms, "cuda")
    builder = VEPBatchBuilder(genome)
    bpe = O.OracleBPE()
    chrom = chroms["chr1"]
    sd = random_init.make_state_dict(CFG, HP, seed=5)
    model = Seq2GenePredictorCombinedModulator(cre_tokenizer=Seq2RegPredictor(**HP), gene_tokenizer=Seq2RegPredictor(**HP), **CFG)
    model.load_state_dict(sd); model.eval().to("cuda"); model.vep = True; attach_trainer(model)
    for g in genes:
        minus = g.strand == "-"
        order = np.argsort(g.cre_start, kind="stable"); order = order[::-1] if minus else order
        k = 3
        a0, a1 = cre_window(g.cre_start[order[k]], g.cre_end[order[k]], 50)
        g0, g1 = gene_window(g.start, g.end, g.strand, 1000, 300000)
        for pos0, alt in ((a0 + 7, None), (g0 + 1234, "ACG"), (a1 + 3, None), (g0 + 4321, None), (g0 - 5000, None)):
            ref = chr(chrom[pos0]).upper()
            alt = alt or [c for c in "ACGT" if c != ref][0]
            v = Variant("chr1", pos0 + 1, ref, alt, tissue=[62, 3])
            het_c = get_iupac_code(ref, alt)
            if g0 < v.pos <= g1 and (het_c == "N" or ref not in "ACGT"):
                # the reference raises here too: encode_with_position rejects a non-IUPAC character (seq.py:93-97)
                with pytest.raises(ValueError, match="invalid character"):
                    builder.build(g, v)
                continue
            batch = builder.build(g, v)
            in_cre = any(cre_window(g.cre_start[i], g.cre_end[i], 50)[0] < v.pos <= cre_window(g.cre_start[i], g.cre_end[i], 50)[1] for i in order)
            in_gene = g0 < v.pos <= g1
            if not in_cre and not in_gene:
                assert batch["variant_type"] == "No overlap" and batch["cre_sequences"] == []
                continue
            for s_idx, rep in enumerate((None, het_c, alt)):
                def window(w0, w1):
                    s = chrom[w0:w1].tobytes().decode()
                    if rep is not None and w0 <= pos0 < w1:
                        s = s[:pos0 - w0] + rep + s[pos0 - w0 + 1:]
                    return O.reverse_complement(s) if minus else s
                toks = np.stack([O.adjust_length(bpe.encode(window(*cre_window(g.cre_start[i], g.cre_end[i], 50))), 200)[0] for i in order])
                assert (batch["cre_sequences"][s_idx][:, 0].cpu().numpy() == toks).all(), f"CRE tokens sample {s_idx}"
                gseq = window(g0, g1)
                gt, gm = O.chunkify(bpe.encode(gseq), 200, 200)
                assert (batch["gene_embeddings"][s_idx][:, 0].cpu().numpy() == gt).all()
                assert (batch["gene_attention_masks"][s_idx][:, 0].cpu().numpy() == gm).all()
                if in_gene:
                    p = pos0 - g0
                    p = len(gseq) - p - 1 if minus else p
                    want = min(bpe.token_at(gseq, p) // 200, 199)
                    assert int(batch["gene_token_position"][s_idx, 0]) == want
            if in_cre:
                hit = [kk for kk, i in enumerate(order) if cre_window(g.cre_start[i], g.cre_end[i], 50)[0] < v.pos <= cre_window(g.cre_start[i], g.cre_end[i], 50)[1]][0]
                assert int(batch["cre_token_position"][0, 0]) == hit
            out = model.predict_step(batch, 0)
            assert len(out["pred_gene_exp"]) == 3 and out["pred_gene_exp"][0].shape == (2, 1)
            assert out["embd"][1].shape == (2, CFG["emb_dim"]) and out["variant_type"] == batch["variant_type"]
            # the three samples of the triplet differ only where the variant falls: ref vs hom predictions differ
            assert np.isfinite(out["pred_gene_exp"][2]).all()


@pytest.mark.parametrize("strand", ["+", "-"])
def test_vep_overlapping_cre_windows_keep_the_last_hit(strand):
    """cCRE windows are +-50 bp wide, so neighbours overlap: the reference's row loop has no break after a hit
    (vepdataset.py:367-407) — every overlapping window is re-encoded and cre_token_position is the LAST one in batch
    order.  On the minus strand its early break (`end_cre < pos`) can stop before a nested window: that one keeps
    the background sequence."""
    from variantformer_b200.datasets.vepdataset import VEPBatchBuilder, Variant
    rng = np.random.default_rng(5)
    chrom = synth.make_chromosome(rng, 700_000, n_rate=0.0)
    genome = Genome.from_arrays({"chr1": chrom}, "cuda")
    cs = np.array([400_000, 400_120, 400_900, 401_000, 401_700]); ce = np.array([400_100, 400_300, 401_600, 401_100, 401_900])
    g = GeneSpec("chr1", 395_000, 420_000, strand, cs, ce, np.arange(5) % 9, [62])
    builder = VEPBatchBuilder(genome)
    bpe = O.OracleBPE()
    minus = strand == "-"
    order = np.argsort(cs, kind="stable"); order = order[::-1] if minus else order
    wins = [cre_window(cs[i], ce[i], 50) for i in order]

    def check(pos0, want_hit, want_applied):
        ref = chr(chrom[pos0]).upper(); alt = [c for c in "ACGT" if c != ref][0]
        batch = builder.build(g, Variant("chr1", pos0 + 1, ref, alt, tissue=[62]))
        assert int(batch["cre_token_position"][0, 0]) == want_hit
        for k, (w0, w1) in enumerate(wins):
            s = chrom[w0:w1].tobytes().decode()
            if k in want_applied:
                s = s[:pos0 - w0] + alt + s[pos0 - w0 + 1:]
            s = O.reverse_complement(s) if minus else s
            tok = O.adjust_length(bpe.encode(s), 200)[0]
            assert (batch["cre_sequences"][2][k, 0].cpu().numpy() == tok).all(), f"hom tokens of window {k}"
    # windows 0 and 1 of the + order overlap on [400070, 400150): both get the variant, the later row wins
    if not minus:
        check(400_100, 1, {0, 1})
    else:
        check(400_100, 4, {3, 4})
        # nested windows on the minus strand: rows by descending start are [401650,401950) [400950,401150)
        # [400850,401650) ...: a variant at 401300 lies in the third row only, but the walk stops at the second row
        # (its end 401150 < pos), so the reference never reaches it: no CRE hit, the window stays unmodified
        ref = chr(chrom[401_300]).upper(); alt = [c for c in "ACGT" if c != ref][0]
        batch = builder.build(g, Variant("chr1", 401_301, ref, alt, tissue=[62]))
        assert bool(torch.isnan(batch["cre_token_position"]).all()) and batch["variant_type"] == "Gene overlap only"
        assert torch.equal(batch["cre_sequences"][2], batch["cre_sequences"][0])


def test_variantprocessor_surface(tmp_path):
    from variantformer_b200.processors.variantprocessor import VariantProcessor
    chroms, var, genes = _world(seed=80, n_genes=2, n_cres=20)
    art, rows = _write_artifacts(tmp_path, chroms, var, genes)
    g = genes[0]
    order = np.argsort(g.cre_start, kind="stable")
    pos0 = int(g.cre_start[order[2]]) + 11                      # inside a CRE window
    while chr(chroms["chr1"][pos0]).upper() not in "ACGT":
        pos0 += 1
    ref = chr(chroms["chr1"][pos0]).upper(); alt = [c for c in "ACGT" if c != ref][1]
    var_df = pd.DataFrame({"chr": ["chr1", "chr1"], "pos": [pos0 + 1, 5], "ref": [ref, "A"], "alt": [alt, "C"],
                           "tissue": ["whole blood,thyroid", "whole blood"], "gene_id": [rows[0]["gene_id"]] * 2})
    vp = VariantProcessor("v4_pcg", base_dir=str(tmp_path), model_overrides=dict(CFG, random_init=5, seq2reg_hyper_parameters=HP))
    out = vp.predict(var_df, str(tmp_path / "out"))
    assert list(out.columns) == ["chrom", "pos", "ref", "alt", "genes", "tissues", "variant_type", "population",
                                 "sample_name", "zygosity", "gene_exp", "gene_emb", "gene_token_embedding", "cre_token_embedding"]
    hit = out[out["pos"] == pos0 + 1]
    assert len(hit) == 6 and set(hit["zygosity"]) == {"0", "1", "2"} and set(hit["tissues"]) == {"whole blood", "thyroid"}
    assert hit["variant_type"].iloc[0] in ("CRE overlap only", "Gene and CRE overlap")
    assert np.isfinite(hit["gene_exp"]).all() and hit["gene_emb"].iloc[0].shape == (CFG["emb_dim"],)
    miss = out[out["pos"] == 5]
    assert (miss["variant_type"] == "No overlap").all() and miss["gene_exp"].isna().all()
    wide = vp.format_scores(out)                                   # the reference's pivot (:454-497): misses are dropped
    assert len(wide) == 2 and {"REF_HG38-0-exp", "REF_HG38-1-exp", "REF_HG38-2-exp"} <= set(wide.columns)
    # no SAMPLE column -> the population aggregate needs the 1KG allele-frequency table of the chromosome
    af_dir = tmp_path / "_artifacts" / "1KG_af_hg38_tables"
    os.makedirs(af_dir, exist_ok=True)
    pd.DataFrame({"chr": ["chr1"], "pos": [pos0 + 1], "ref": [ref], "alt": [alt], "AF_EUR": [0.1], "AF_AFR": ["."],
                  "AF_EAS": [0.2], "AF_SAS": [0.0], "AF_AMR": [0.3]}).to_csv(af_dir / "1KG_hg38_af_chr1.tsv", sep="\t", index=False)
    scores = vp.eqtl_scores(wide)
    assert "VF-REF_HG38-2-exp-log2fc" in scores.columns and np.isfinite(scores["VF-REF_HG38-2-exp-log2fc"]).all()
    want = np.log2((wide["REF_HG38-2-exp"] + 1e-10) / (wide["REF_HG38-0-exp"] + 1e-10))
    assert np.allclose(scores["VF-REF_HG38-2-exp-log2fc"], want)
    with pytest.raises(FileExistsError):
        vp.predict(var_df, str(tmp_path / "out"))               # refuses to overwrite, like the reference
