"""Generate tests/golden/model_golden.npz from the REFERENCE's own PyTorch classes.

Run in the build container only:  python tests/golden/make_model_golden.py

A reduced-width model (same architecture flags as configs/vf_model.yaml v4_pcg:
use_context, multi_registry, bigger shared head, ALiBi self-attention, head_dim
48 for seq2gene and 64 for seq2reg) is built from the reference classes
(Seq2RegPredictor x2 + Seq2GenePredictorCombinedModulator), loaded with the
deterministic weights of variantformer_b200.utils.random_init (asserting that
key names and shapes equal the reference's own state_dict), and run through the
reference's unmodified `predict_step` in fp32 on CPU.  Inputs and outputs are
stored; weights are regenerated from the seed by the tests.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import refshim  # noqa: E402
from variantformer_b200.utils import random_init  # noqa: E402

CFG = dict(random_init.V4_PCG_MODEL, emb_dim=192, gene_emb_dim=128, num_heads=4, num_layers=3, token_dim=128)
HP = dict(random_init.SEQ2REG_HP, embedding_dim=128, num_heads=2, num_layers=2)
SEED = 7


def synth_batch(rng, n_genes, c_range, g_range, t_list, max_len=200, vocab=500):
    def windows(n, full):
        tok = np.zeros((n, 1, max_len), np.int64); mask = np.ones((n, 1, max_len), bool)
        for i in range(n):
            L = max_len if (full and i < n - 1) else int(rng.integers(3, max_len + 1))
            tok[i, 0, :L] = rng.integers(4, vocab, L); mask[i, 0, :L] = False
        return torch.from_numpy(tok), torch.from_numpy(mask)
    b = {k: [] for k in ("cre_sequences", "cre_attention_masks", "tissue_context", "cre_labels", "ref_cre_labels",
                         "gene_embeddings", "gene_attention_masks")}
    strands = []
    for g in range(n_genes):
        C = int(rng.integers(*c_range)); G = int(rng.integers(*g_range))
        x, m = windows(C, False); gx, gm = windows(G, True)
        b["cre_sequences"].append(x); b["cre_attention_masks"].append(m)
        b["gene_embeddings"].append(gx); b["gene_attention_masks"].append(gm)
        b["tissue_context"].append(torch.tensor(t_list[g], dtype=torch.long))
        b["ref_cre_labels"].append(torch.from_numpy(rng.integers(0, 9, C)))
        b["cre_labels"].append(torch.zeros(C, dtype=torch.long))
        strands.append([g % 2])
    b["strand_val"] = torch.tensor(strands, dtype=torch.long)
    return b


def main():
    sd = random_init.make_state_dict(CFG, HP, seed=SEED)
    model = refshim.build_reference_model(
        {k: v for k, v in CFG.items()}, HP, seed=0)
    ref_sd = model.state_dict()
    assert set(ref_sd) == set(sd), (sorted(set(ref_sd) ^ set(sd))[:10])
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd)
    rng = np.random.default_rng(99)
    batch = synth_batch(rng, 2, (5, 9), (2, 4), [[62, 0, 14], [7]])
    with torch.no_grad():
        out = model.predict_step(batch, 0)
    save = {"n_genes": 2}
    for g in range(2):
        for k in ("cre_sequences", "cre_attention_masks", "tissue_context", "ref_cre_labels", "gene_embeddings",
                  "gene_attention_masks"):
            save[f"{k}_{g}"] = batch[k][g].numpy()
        save[f"pred_{g}"] = out["pred_gene_exp"][g]; save[f"emb_{g}"] = out["embeddings"][g]
    save["strand_val"] = batch["strand_val"].numpy()
    # the full-size key/shape contract, stored as text so the CPU suite can assert it without the reference
    full = refshim.build_reference_model(dict(random_init.V4_PCG_MODEL, num_layers=2), dict(random_init.SEQ2REG_HP, num_layers=1))
    save["full_keys"] = np.array(sorted(f"{k}:{tuple(v.shape)}" for k, v in full.state_dict().items()))
    path = os.path.join(HERE, "model_golden.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, os.path.getsize(path), "bytes; preds", [p.ravel().tolist() for p in out["pred_gene_exp"]])

    # second, larger fixture from the same reference classes and weights: multi-tile CRE attention (C = 300 > 128 and
    # > 256), G = 130 gene chunks (> 128: two query tiles per tissue copy), both strand flags, T = 5 and the full
    # T = 63 tissue axis.  Tokens are stored as int16, masks as uint8 (tests/common.py:load_model_golden widens them).
    rng = np.random.default_rng(100)
    big = synth_batch(rng, 2, (300, 301), (130, 131), [[62, 0, 14, 33, 5], list(range(63))])
    big2 = synth_batch(rng, 1, (150, 151), (40, 41), [list(range(63))])
    for k in big:                                         # gene 1 of the fixture = the smaller C/G gene with T = 63
        if k != "strand_val":
            big[k][1] = big2[k][0]
    with torch.no_grad():
        out = model.predict_step(big, 0)
    save = {"n_genes": 2}
    for g in range(2):
        for k in ("cre_sequences", "gene_embeddings"):
            save[f"{k}_{g}"] = big[k][g].numpy().astype(np.int16)
        for k in ("cre_attention_masks", "gene_attention_masks"):
            save[f"{k}_{g}"] = big[k][g].numpy().astype(np.uint8)
        for k in ("tissue_context", "ref_cre_labels"):
            save[f"{k}_{g}"] = big[k][g].numpy().astype(np.int16)
        save[f"pred_{g}"] = out["pred_gene_exp"][g]; save[f"emb_{g}"] = out["embeddings"][g]
    save["strand_val"] = big["strand_val"].numpy()
    path = os.path.join(HERE, "model_golden_large.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, os.path.getsize(path), "bytes; shapes", [e.shape for e in out["embeddings"]])


if __name__ == "__main__":
    main()
