"""Small host helpers mirroring the hot-path parts of the reference's utils/functions.py."""
import os

import numpy as np
import torch

_COMP = str.maketrans("ACGTRYKMBVDHacgtrykmbvdh", "TGCAYRMKVBHDtgcayrmkvbhd")


def precision2dtype(precision_str: str) -> torch.dtype:
    """Lightning precision string -> dtype (utils/functions.py:12-32)."""
    p = precision_str.lower().strip()
    if "bf16" in p:
        return torch.bfloat16
    if "16" in p:
        return torch.float16
    if "32" in p:
        return torch.float32
    raise ValueError(f"Unknown precision string: {precision_str}")


def reverse_complement(sequence: str) -> str:
    """Host-side convenience (utils/functions.py:129-172); the batched path does this inside vf_encode_windows."""
    return sequence[::-1].translate(_COMP)


# ---- variant-effect scores (utils/functions.py:178-301) -----------------------------------------------------------
_POPS = ("AFR", "AMR", "EAS", "EUR", "SAS")


def merge_pop_stat(df, af_path):
    """Left-join the 1000-Genomes allele-frequency tables (`1KG_hg38_af_<chr>.tsv`, columns chr/pos/ref/alt/AF_*) onto
    the score table (utils/functions.py:178-204)."""
    import pandas as pd
    parts = []
    for c in df["chr"].unique():
        af = pd.read_csv(os.path.join(af_path, f"1KG_hg38_af_{c}.tsv"), sep="\t")
        parts.append(df[df["chr"] == c].merge(af, on=["chr", "pos", "ref", "alt"], how="left").reset_index(drop=True))
    out = pd.concat(parts, ignore_index=True)
    for pop in _POPS:
        out["AF_" + pop] = out["AF_" + pop].replace(".", np.nan).astype(float)
    return out


def gene_pop_agg_score(df, score_cols, score_type="log2fc"):
    """Allele-frequency weighted mean of the per-population homozygous scores (utils/functions.py:207-247), for the whole
    table at once: rows whose scores are all NaN give NaN, rows whose valid frequencies sum to 0 fall back to the plain
    mean of the valid scores."""
    if f"VF-REF_HG38-2-exp-{score_type}" in score_cols:
        score_cols = [c for c in score_cols if "REF_HG38-2" not in c]
    pop_cols = [c for c in score_cols if any(c.startswith(f"VF-{p}-2") for p in _POPS)]
    af_cols = ["AF_" + c.split("-")[1] for c in pop_cols]
    sc = df[score_cols].to_numpy(float)
    af = df[af_cols].to_numpy(float)
    if sc.shape[1] != af.shape[1]:                       # (the reference pairs score and frequency columns by position)
        raise ValueError("every score column needs a population allele-frequency column")
    valid = ~np.isnan(sc)
    w = np.where(valid, af, 0.0)
    wsum = w.sum(1)                                      # NaN when a valid score has no frequency ('.'): plain mean then,
    n_valid = valid.sum(1)                               # exactly like the reference's `np.sum(valid_af) > 0` test
    with np.errstate(invalid="ignore", divide="ignore"):
        weighted = (np.where(valid, sc, 0.0) * w).sum(1) / wsum
        plain = np.where(valid, sc, 0.0).sum(1) / n_valid
    agg = np.where(n_valid == 0, np.nan, np.where(wsum > 0, weighted, plain))
    df = df.copy()
    df["VF-agg-" + score_type + "-weighted"] = agg
    return df


def generate_log2fc_score(df, af_path=None):
    """log2((pop + 1e-10) / (ref + 1e-10)) of every homozygous-ALT expression column against `REF_HG38-0-exp`
    (utils/functions.py:250-301).  Without a `SAMPLE-2-exp` column the population scores are aggregated with the
    1000-Genomes allele frequencies found under `af_path`."""
    ref_col = "REF_HG38-0-exp"
    pop_columns = [c for c in df.columns if any(c.startswith(p + "-2") for p in _POPS + ("REF_HG38", "SAMPLE"))]
    keys = ["variant_id", "genes", "tissues", "ref", "alt", "chr", "pos"]
    df = df[[ref_col] + pop_columns + keys].reset_index(drop=True).copy()
    ref = df[ref_col].to_numpy(float)
    score_cols = []
    for c in pop_columns:
        df["VF-" + c + "-log2fc"] = np.log2((df[c].to_numpy(float) + 1e-10) / (ref + 1e-10))
        score_cols.append("VF-" + c + "-log2fc")
    if not any(c.startswith("SAMPLE-2") for c in pop_columns):
        df = gene_pop_agg_score(merge_pop_stat(df, af_path), score_cols, score_type="log2fc")
        return df[keys + ["VF-agg-log2fc-weighted"] + score_cols]
    return df[keys + score_cols]
