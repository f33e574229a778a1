"""Shared helpers for parity tests: golden loading, tolerances, synthetic batches."""
import os

import numpy as np
import torch

from variantformer_b200.utils import random_init

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GOLD_CFG = dict(random_init.V4_PCG_MODEL, emb_dim=192, gene_emb_dim=128, num_heads=4, num_layers=3, token_dim=128)
GOLD_HP = dict(random_init.SEQ2REG_HP, embedding_dim=128, num_heads=2, num_layers=2)
GOLD_SEED = 7


def load_model_golden():
    z = np.load(os.path.join(GOLDEN, "model_golden.npz"), allow_pickle=True)
    n = int(z["n_genes"])
    batch = {k: [torch.from_numpy(z[f"{k}_{g}"]) for g in range(n)]
             for k in ("cre_sequences", "cre_attention_masks", "tissue_context", "ref_cre_labels",
                       "gene_embeddings", "gene_attention_masks")}
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


def rel_err(got, want):
    """max |got-want| / max |want|  — 'max rel. error' normalised by the tensor's scale."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def pearson(got, want):
    got = np.asarray(got, np.float64).ravel(); want = np.asarray(want, np.float64).ravel()
    if got.size < 2 or want.std() == 0:
        return 1.0
    return float(np.corrcoef(got, want)[0, 1])
