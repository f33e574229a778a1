"""Variant-effect scores (VariantProcessor.eqtl_scores -> utils.functions.generate_log2fc_score) against a fixture the
reference's own utils/functions.py produced (tests/golden/make_scores_golden.py)."""
import json
import os

import numpy as np
import pandas as pd

from variantformer_b200.utils.functions import generate_log2fc_score

GOLD = os.path.join(os.path.dirname(__file__), "golden", "scores_golden.json")


def _frame(d):
    return pd.DataFrame({k: [np.nan if v is None else v for v in col] for k, col in d.items()})


def _same(a, b):
    assert list(a.columns) == list(b.columns)
    for c in a.columns:
        if a[c].dtype.kind == "f" or b[c].dtype.kind == "f":
            assert np.allclose(a[c].to_numpy(float), b[c].to_numpy(float), equal_nan=True, rtol=1e-12, atol=0), c
        else:
            assert (a[c].to_numpy() == b[c].to_numpy()).all(), c


def test_log2fc_and_af_weighted_aggregate_match_the_reference(tmp_path):
    g = json.load(open(GOLD))
    for c, t in g["af"].items():
        pd.DataFrame(t).to_csv(tmp_path / f"1KG_hg38_af_{c}.tsv", sep="\t", index=False)
    df = _frame(g["input"])
    _same(generate_log2fc_score(df.copy(), str(tmp_path)), _frame(g["want_population"]))
    df["SAMPLE-2-exp"] = g["sample_col"]
    _same(generate_log2fc_score(df.copy(), str(tmp_path)), _frame(g["want_sample"]))
