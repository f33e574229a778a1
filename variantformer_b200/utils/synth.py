"""Seeded synthetic genome / variants / gene+CRE layout (SURVEY.md §8(d), config 1 stand-in for the
example VCF, which cannot be downloaded offline).  Pure numpy; used by tests and bench.py."""
import numpy as np


def make_chromosome(rng, length, softmask_frac=0.3, n_rate=1e-4):
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), length)
    # soft-masked (lower-case) runs covering ~softmask_frac of the sequence
    n_runs = max(1, int(length * softmask_frac / 300))
    starts = rng.integers(0, length, n_runs); lens = rng.integers(50, 550, n_runs)
    for s, l in zip(starts, lens):
        seq[s:s + l] |= 0x20
    n_n = rng.binomial(length, n_rate)
    for s in rng.integers(0, length, n_n):
        seq[s:s + int(rng.integers(1, 51))] = ord("N")
    return seq


def make_variants(rng, chrom_seq, rate=1 / 700, indel_frac=0.1):
    """-> dict(pos, ref_len, alt(list of bytes), gt) sorted by position; het:hom = 2:1."""
    length = len(chrom_seq)
    n = rng.binomial(length, rate)
    pos = np.unique(rng.integers(0, length - 8, n))
    ref_len = np.ones(len(pos), np.int32); alt = []; gt = rng.choice(np.array([1, 1, 2], np.uint8), len(pos))
    bases = b"ACGT"
    for i, p in enumerate(pos):
        r = chr(chrom_seq[p]).upper()
        if rng.random() < indel_frac:
            k = int(rng.integers(1, 6))
            if rng.random() < 0.5:                      # insertion: REF base + k random bases
                alt.append((r + "".join(chr(bases[j]) for j in rng.integers(0, 4, k))).encode())
            else:                                       # deletion of k bases after the anchor
                ref_len[i] = 1 + k; alt.append(r.encode())
        else:
            choices = [c for c in "ACGT" if c != r]
            alt.append(choices[int(rng.integers(0, len(choices)))].encode())
    return dict(pos=pos.astype(np.int64), ref_len=ref_len, alt=alt, gt=gt)


def make_gene_layout(rng, chrom_len, n_cres, strand=None, body_len=None, cre_span=600_000):
    """One gene + its CRE table on a chromosome.  -> dict(start, end, strand, cre_start[], cre_end[], labels[])."""
    if body_len is None:
        body_len = int(np.clip(rng.lognormal(np.log(25_000), 1.0), 1_000, 2_300_000))
    strand = strand or ("+" if rng.random() < 0.5 else "-")
    margin = 310_000 + cre_span
    start = int(rng.integers(margin, max(margin + 1, chrom_len - margin - body_len)))
    end = start + body_len
    centre = (start + end) // 2
    cs = np.sort(rng.integers(max(0, centre - cre_span), min(chrom_len - 500, centre + cre_span), n_cres))
    ce = cs + rng.integers(150, 351, n_cres)
    return dict(start=start, end=end, strand=strand, cre_start=cs.astype(np.int64), cre_end=ce.astype(np.int64),
                labels=rng.integers(0, 9, n_cres).astype(np.int64))


def token_batch(seed, n_genes, C, G, T, max_len=200, vocab=500, mean_cre_tokens=97, tissues=None):
    """Token-level synthetic batch (reference collate keys) at benchmark shapes: C CRE windows of ~97 valid tokens,
    G full gene chunks (last one ragged), T tissues per gene."""
    import torch
    rng = np.random.default_rng(seed)
    b = {k: [] for k in ("cre_sequences", "cre_attention_masks", "tissue_context", "cre_labels", "ref_cre_labels",
                         "gene_embeddings", "gene_attention_masks")}
    for g in range(n_genes):
        lens = np.clip(rng.normal(mean_cre_tokens, 15, C).astype(int), 8, max_len)
        tok = rng.integers(4, vocab, (C, 1, max_len)).astype(np.int64)
        mask = np.arange(max_len)[None, None, :] >= lens[:, None, None]
        tok[mask] = 0
        gt = rng.integers(4, vocab, (G, 1, max_len)).astype(np.int64); gm = np.zeros((G, 1, max_len), bool)
        last = int(rng.integers(1, max_len + 1)); gt[-1, 0, last:] = 0; gm[-1, 0, last:] = True
        b["cre_sequences"].append(torch.from_numpy(tok)); b["cre_attention_masks"].append(torch.from_numpy(mask))
        b["gene_embeddings"].append(torch.from_numpy(gt)); b["gene_attention_masks"].append(torch.from_numpy(gm))
        tis = tissues[g] if tissues is not None else list(range(T))
        b["tissue_context"].append(torch.tensor(tis, dtype=torch.long))
        b["ref_cre_labels"].append(torch.from_numpy(rng.integers(0, 9, C)))
        b["cre_labels"].append(torch.zeros(C, dtype=torch.long))
    b["strand_val"] = torch.zeros(n_genes, 1, dtype=torch.long)
    return b
