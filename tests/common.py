"""Shared helpers for parity tests: golden loading, tolerances, synthetic batches."""
import os

import numpy as np
import torch

from variantformer_b200.utils import random_init

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GOLD_CFG = dict(random_init.V4_PCG_MODEL, emb_dim=192, gene_emb_dim=128, num_heads=4, num_layers=3, token_dim=128)
GOLD_HP = dict(random_init.SEQ2REG_HP, embedding_dim=128, num_heads=2, num_layers=2)
GOLD_SEED = 7


def load_model_golden(name="model_golden.npz"):
    """Fixtures written by tests/golden/make_model_golden.py from the reference's own classes: "model_golden.npz" (tiny:
    5-8 CREs, 2-3 chunks) and "model_golden_large.npz" (C = 300 / 150, G = 130 / 40, T = 5 / 63, both strand flags)."""
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=True)
    n = int(z["n_genes"])
    widen = {"cre_sequences": np.int64, "gene_embeddings": np.int64, "tissue_context": np.int64, "ref_cre_labels": np.int64,
             "cre_attention_masks": bool, "gene_attention_masks": bool}
    batch = {k: [torch.from_numpy(z[f"{k}_{g}"].astype(dt)) for g in range(n)] for k, dt in widen.items()}
    batch["strand_val"] = torch.from_numpy(z["strand_val"])
    want = {"pred_gene_exp": [z[f"pred_{g}"] for g in range(n)], "embeddings": [z[f"emb_{g}"] for g in range(n)]}
    return batch, want, z


def synth_batch(seed, n_genes, C, G, tissues, max_len=200, vocab=500, mean_cre_tokens=97):
    """Synthetic token-level batch with the reference's collate keys (vcfdataset.py:53-63)."""
    rng = np.random.default_rng(seed)
    b = {k: [] for k in ("cre_sequences", "cre_attention_masks", "tissue_context", "cre_labels", "ref_cre_labels",
                         "gene_embeddings", "gene_attention_masks")}
    for g in range(n_genes):
        c = C[g] if isinstance(C, (list, tuple)) else C
        gg = G[g] if isinstance(G, (list, tuple)) else G
        tok = np.zeros((c, 1, max_len), np.int64); mask = np.ones((c, 1, max_len), bool)
        lens = np.clip(rng.normal(mean_cre_tokens, 15, c).astype(int), 8, max_len)
        for i in range(c):
            tok[i, 0, :lens[i]] = rng.integers(4, vocab, lens[i]); mask[i, 0, :lens[i]] = False
        gt = rng.integers(4, vocab, (gg, 1, max_len)).astype(np.int64); gm = np.zeros((gg, 1, max_len), bool)
        last = int(rng.integers(1, max_len + 1)); gt[-1, 0, last:] = 0; gm[-1, 0, last:] = True
        b["cre_sequences"].append(torch.from_numpy(tok)); b["cre_attention_masks"].append(torch.from_numpy(mask))
        b["gene_embeddings"].append(torch.from_numpy(gt)); b["gene_attention_masks"].append(torch.from_numpy(gm))
        b["tissue_context"].append(torch.tensor(tissues[g], dtype=torch.long))
        b["ref_cre_labels"].append(torch.from_numpy(rng.integers(0, 9, c)))
        b["cre_labels"].append(torch.zeros(c, dtype=torch.long))
    b["strand_val"] = torch.zeros(n_genes, 1, dtype=torch.long)
    return b


# ---- tolerance of the floating-point parity tests -------------------------------------------------------------------
# north_star: "max rel. error <= 1e-2, Pearson >= 0.9999" against the reference's forward on the same weights and inputs.
# SURVEY section 8(d) suggests a per-element form |a-b| / max(|b|, eps) with eps = 1e-3 RMS(b).  THAT FORM IS NOT USED
# HERE, DELIBERATELY: a bf16 forward cannot meet it — an embedding entry near zero (|b| ~ 1e-3 RMS) carries the same
# absolute rounding noise as its neighbours (~1e-3 of the tensor scale), i.e. a "relative" error of order 1; the
# reference's own bf16-mixed forward scores ~0.3 on that metric against its fp32 self.  What is asserted instead:
#   rel_err      = max |got - want| / max |want|          <= 1e-2   (one scale per tensor)
#   pearson(got, want)                                     >= 0.9999
# and, reported next to it and asserted at the looser bound that bf16 supports,
#   rel_err_rms  = max |got - want| / RMS(want)            <= 5e-2   (3-4x stricter than rel_err on these tensors)
# The reference's own end-to-end tolerances are absolute (tests/test_vep.py:216-257 atol 1e-3 on log2FC; :389-403 atol
# 0.1 on expression, 1 on embeddings) and far looser than either.
REL_ERR_MAX, PEARSON_MIN, REL_ERR_RMS_MAX = 1e-2, 0.9999, 5e-2


def rel_err(got, want):
    """max |got-want| / max |want|  — 'max rel. error' normalised by the tensor's scale."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def rel_err_rms(got, want):
    """max |got-want| / RMS(want): the stricter normalisation (see the note above)."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.sqrt(np.mean(want * want)), 1e-30))


def parity_report(got, want):
    """-> dict of the three figures + `ok` against the stated bounds."""
    r = {"rel_err_max_norm": rel_err(got, want), "rel_err_rms_norm": rel_err_rms(got, want), "pearson": pearson(got, want)}
    r["ok"] = bool(r["rel_err_max_norm"] <= REL_ERR_MAX and r["pearson"] >= PEARSON_MIN and
                   r["rel_err_rms_norm"] <= REL_ERR_RMS_MAX)
    return r


def pearson(got, want):
    got = np.asarray(got, np.float64).ravel(); want = np.asarray(want, np.float64).ravel()
    if got.size < 2 or want.std() == 0:
        return 1.0
    return float(np.corrcoef(got, want)[0, 1])
