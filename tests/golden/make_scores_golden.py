"""Generate tests/golden/scores_golden.json from the REFERENCE's own utils/functions.py (generate_log2fc_score with the
allele-frequency weighted aggregate).  Run in the build container only:  python tests/golden/make_scores_golden.py"""
import importlib.util
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("pybedtools", types.ModuleType("pybedtools"))          # utils/functions.py:1 imports it
sys.path.insert(0, "/root/reference")
spec = importlib.util.spec_from_file_location("ref_functions", "/root/reference/utils/functions.py")
R = importlib.util.module_from_spec(spec); spec.loader.exec_module(R)
warnings.filterwarnings("ignore")
POPS = ("AFR", "AMR", "EAS", "EUR", "SAS")


def main():
    rng = np.random.default_rng(0)
    n = 40
    df = pd.DataFrame({"variant_id": [f"v{i}" for i in range(n)], "genes": "g", "tissues": "t", "ref": "A", "alt": "C",
                       "chr": ["chr1"] * 20 + ["chr2"] * 20, "pos": np.arange(n) + 1, "REF_HG38-0-exp": rng.random(n) + 0.1})
    for p in POPS + ("REF_HG38",):
        col = rng.random(n) + 0.1
        if p != "REF_HG38":
            col[rng.random(n) < 0.3] = np.nan
        df[p + "-2-exp"] = col
    df.loc[3, [p + "-2-exp" for p in POPS]] = np.nan                          # a row without any population score
    d = tempfile.mkdtemp()
    af_tables = {}
    for c in ("chr1", "chr2"):
        sub = df[df.chr == c]
        af = pd.DataFrame({"chr": c, "pos": sub.pos, "ref": "A", "alt": "C",
                           **{"AF_" + p: np.round(rng.random(len(sub)), 6) for p in ("EUR", "AFR", "EAS", "SAS", "AMR")}})
        af.loc[af.index[2], ["AF_EUR", "AF_AFR", "AF_EAS", "AF_SAS", "AF_AMR"]] = 0.0      # all-zero frequencies
        af["AF_AFR"] = af["AF_AFR"].astype(object); af.loc[af.index[5], "AF_AFR"] = "."     # missing frequency
        af.to_csv(os.path.join(d, f"1KG_hg38_af_{c}.tsv"), sep="\t", index=False)
        af_tables[c] = af.astype(object).to_dict(orient="list")
    pop = R.generate_log2fc_score(df.copy(), d)
    df2 = df.copy(); df2["SAMPLE-2-exp"] = rng.random(n)
    smp = R.generate_log2fc_score(df2.copy(), d)
    ser = lambda t: {c: [None if (isinstance(v, float) and np.isnan(v)) else v for v in t[c].tolist()] for c in t.columns}
    json.dump({"input": ser(df), "sample_col": df2["SAMPLE-2-exp"].tolist(), "af": af_tables,
               "want_population": ser(pop), "want_sample": ser(smp)},
              open(os.path.join(HERE, "scores_golden.json"), "w"))
    print("wrote scores_golden.json", pop.shape, smp.shape)


if __name__ == "__main__":
    main()
