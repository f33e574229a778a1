"""AD-risk head — the reference's `ADrisk` / `ADriskFromVCF` call shapes (processors/ad_risk.py:20-66, 69-176) on the GPU.

The reference keeps one treelite GBDT per (gene, tissue) (S3 artifacts) and evaluates them one row at a time with
`treelite.gtil.predict` inside a pandas loop.  Here all forests of a run live in ONE device-resident bank
(`ForestBank`: concatenated node arrays) and every row of embeddings is scored by `vf_forest_predict` in a single launch,
each row picking its own forest.

Model format (`*.npz`, one file per forest or one bank): tree_root int32 [n_trees], feat int32 [n_nodes] (-1 = leaf; bit 31
on a split = missing values go left), thr fp32, left / right int32 (node indices inside the forest), value fp32 (leaf
contribution with the learning rate folded in), base fp32 (raw-score offset), op_lt (0: x <= thr goes left).
Converters: `forest_from_sklearn` (GradientBoostingClassifier, binary) and `forest_from_treelite_json`
(`treelite.Model.dump_as_json()` of a binary-logistic model) — treelite itself is not needed at inference time; its
serialised `.tl` checkpoints must be dumped to JSON once where treelite is installed.
"""
import json
import os

import numpy as np
import pandas as pd
import torch

from .. import ops

VF_DIMS = 1536
_KEYS = ("tree_root", "feat", "thr", "left", "right", "value")


def forest_from_sklearn(gbc):
    """sklearn.ensemble.GradientBoostingClassifier (binary) -> forest dict (numpy)."""
    assert gbc.n_classes_ == 2, "binary classifiers only (P(AD) vs P(no AD))"
    roots, feat, thr, left, right, value, off = [], [], [], [], [], [], 0
    for est in gbc.estimators_[:, 0]:
        t = est.tree_
        roots.append(off)
        leaf = t.children_left == -1
        feat.append(np.where(leaf, -1, t.feature).astype(np.int32))
        # x (fp32) <= thr (fp64)  <=>  x <= the largest fp32 that does not exceed thr: round the threshold DOWN
        t32 = t.threshold.astype(np.float32)
        thr.append(np.where(t32.astype(np.float64) > t.threshold, np.nextafter(t32, np.float32(-np.inf)), t32).astype(np.float32))
        left.append(np.where(leaf, 0, t.children_left + off).astype(np.int32))
        right.append(np.where(leaf, 0, t.children_right + off).astype(np.int32))
        value.append((gbc.learning_rate * t.value[:, 0, 0]).astype(np.float32))
        off += t.node_count
    # the prior of the default init estimator: log-odds of the positive class
    p1 = float(gbc.init_.class_prior_[1]) if hasattr(gbc.init_, "class_prior_") else 0.5
    base = np.float32(np.log(p1 / (1.0 - p1)))
    return dict(tree_root=np.asarray(roots, np.int32), feat=np.concatenate(feat), thr=np.concatenate(thr),
                left=np.concatenate(left), right=np.concatenate(right), value=np.concatenate(value), base=base, op_lt=0)


def forest_from_treelite_json(text):
    """`treelite.Model.dump_as_json()` of a binary classifier with a sigmoid post-processor -> forest dict (numpy)."""
    m = json.loads(text) if isinstance(text, str) else text
    roots, feat, thr, left, right, value, off = [], [], [], [], [], [], 0
    op = None
    for tree in m["trees"]:
        nodes = tree["nodes"]
        ids = {n["node_id"]: i for i, n in enumerate(nodes)}
        roots.append(off + ids[tree.get("root_id", 0)])
        for n in nodes:
            if "leaf_value" in n:
                lv = n["leaf_value"]
                feat.append(-1); thr.append(0.0); left.append(0); right.append(0)
                value.append(float(lv[0] if isinstance(lv, list) else lv))
            else:
                op = n.get("comparison_op", "<=") if op is None else op
                assert n.get("comparison_op", op) == op, "mixed comparison operators in one model"
                f = int(n["split_feature_id"]) | (0x80000000 if n.get("default_left", False) else 0)
                feat.append(np.int32(np.uint32(f))); thr.append(float(n["threshold"]))
                left.append(off + ids[n["left_child"]]); right.append(off + ids[n["right_child"]]); value.append(0.0)
        off += len(nodes)
    base = m.get("base_scores", [m.get("global_bias", 0.0)])
    base = float(base[0] if isinstance(base, list) else base)
    assert op in (None, "<=", "<"), f"unsupported comparison operator {op}"
    return dict(tree_root=np.asarray(roots, np.int32), feat=np.asarray(feat, np.int32), thr=np.asarray(thr, np.float32),
                left=np.asarray(left, np.int32), right=np.asarray(right, np.int32), value=np.asarray(value, np.float32),
                base=np.float32(base), op_lt=int(op == "<"))


def save_forest(path, forest):
    np.savez_compressed(path, **forest)


def load_forest(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


class ForestBank:
    """Many forests in one set of device arrays; rows are scored together, each under its own forest."""

    def __init__(self, forests, device="cuda"):
        forests = list(forests)
        assert forests and len({int(f.get("op_lt", 0)) for f in forests}) == 1, "one comparison convention per bank"
        node_off, tree_off = 0, [0]
        cat = {k: [] for k in _KEYS}
        for f in forests:
            cat["tree_root"].append(np.asarray(f["tree_root"], np.int32) + node_off)
            leaf = np.asarray(f["feat"], np.int32) == -1
            cat["feat"].append(np.asarray(f["feat"], np.int32)); cat["thr"].append(np.asarray(f["thr"], np.float32))
            cat["left"].append(np.where(leaf, 0, np.asarray(f["left"], np.int32) + node_off).astype(np.int32))
            cat["right"].append(np.where(leaf, 0, np.asarray(f["right"], np.int32) + node_off).astype(np.int32))
            cat["value"].append(np.asarray(f["value"], np.float32))
            node_off += len(f["feat"]); tree_off.append(tree_off[-1] + len(f["tree_root"]))
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.dev = {k: t(np.concatenate(v)) for k, v in cat.items()}
        self.dev["tree_off"] = t(np.asarray(tree_off, np.int32))
        self.dev["base"] = t(np.asarray([float(f["base"]) for f in forests], np.float32))
        self.dev["op_lt"] = int(forests[0].get("op_lt", 0))
        self.n_forests, self.device = len(forests), torch.device(device)

    def predict_proba(self, x, forest_ids):
        """x [n, d] fp32 (numpy or tensor), forest_ids [n] -> P(class 1) numpy [n]."""
        x = torch.as_tensor(np.asarray(x, np.float32) if not torch.is_tensor(x) else x, dtype=torch.float32,
                            device=self.device).contiguous()
        ids = torch.as_tensor(np.asarray(forest_ids, np.int32), device=self.device)
        assert x.dim() == 2 and ids.numel() == x.shape[0] and int(ids.max()) < self.n_forests
        return ops.forest_predict(x, ids, self.dev).cpu().numpy()


class LocalPredictorManifest:
    """`GeneTissueManifestLookup` stand-in (utils/assets.py is S3 plumbing, out of scope): <dir>/<gene_id>_<tissue_id>.npz."""

    def __init__(self, directory):
        self.directory = directory

    def get_file_path(self, gene_id, tissue_id):
        p = os.path.join(self.directory, f"{gene_id}_{int(tissue_id)}.npz")
        return p if os.path.exists(p) else None


class ADrisk:
    """processors/ad_risk.py:20-66: one (gene, tissue) predictor; `__call__(embeddings [n, 1536]) -> P(AD) [n]`."""

    def __init__(self, gene_id: str, tissue_id: int, model_class: str = "v4_pcg", manifest=None):
        assert model_class in ["v4_ag", "v4_pcg"], "model_class should be either 'v4_ag' or 'v4_pcg'"
        assert type(tissue_id) is int, "tissue_id should be an integer"
        assert type(gene_id) is str, "gene_id should be a string"
        self.gene_id, self.tissue_id, self.model_class = gene_id, tissue_id, model_class
        self.gene_tissue_manifest = manifest
        self.ad_preds = self._load_ad_predictor()

    def _load_ad_predictor(self):
        fname = self.gene_tissue_manifest.get_file_path(self.gene_id, self.tissue_id) if self.gene_tissue_manifest else None
        if fname is None:
            raise FileNotFoundError(f"AD predictor not found for gene {self.gene_id} and tissue {self.tissue_id}")
        self.predictor = ForestBank([load_forest(fname)])
        return self.predictor

    def __call__(self, gene_tissue_embeds: np.ndarray):
        return self.predictor.predict_proba(gene_tissue_embeds, np.zeros(len(gene_tissue_embeds), np.int32))


class ADriskFromVCF:
    """processors/ad_risk.py:69-176: VCF -> embeddings (VCFProcessor) -> per-(gene, tissue) AD risk.  The reference's
    pandas row loop with one treelite deserialisation + prediction per row becomes one forest-bank launch."""

    def __init__(self, model_class: str = "v4_pcg", vcf_processor=None, manifest=None):
        from .vcfprocessor import VCFProcessor
        self.model_class = model_class
        self.ad_preds = manifest
        self.vcf_processor = vcf_processor or VCFProcessor(model_class=model_class)
        tissues = self.vcf_processor.tissue_vocab
        self.tissue_map = pd.DataFrame({"tissue": list(tissues.keys())},
                                       index=pd.Index(list(tissues.values()), name="tissue_id"))
        self.genes_map = self.vcf_processor.get_genes().set_index("gene_id")
        self.model, self.checkpoint_path, self.trainer = self.vcf_processor.load_model()

    def _format_query(self, gene_ids, tissue_ids):
        assert len(gene_ids) == len(tissue_ids), \
            "Please map gene ids to tissue ids, there should be 2 lists of the same length"
        for gene_id, tissue_id in zip(gene_ids, tissue_ids):
            assert type(tissue_id) is int, "tissue_id should be an integer"
            assert type(gene_id) is str, "gene_id should be a string"
        return pd.DataFrame({"gene_id": gene_ids, "tissues": self.tissue_map.loc[tissue_ids]["tissue"].tolist()})

    def __call__(self, vcf_file: str, gene_ids, tissue_ids) -> pd.DataFrame:
        ds, dl = self.vcf_processor.create_data(vcf_file, self._format_query(gene_ids, tissue_ids))
        df = self.vcf_processor.predict(self.model, self.checkpoint_path, self.trainer, dl, ds)
        df = df.copy()
        df["tissue_id"] = list(tissue_ids)
        df["embedding"] = df["embeddings"].apply(lambda e: np.asarray(e)[0])
        df["predicted_expression"] = df["predicted_expression"].apply(lambda x: float(np.asarray(x).ravel()[0]))
        return self._predict_ad_risk(df)

    def _predict_ad_risk(self, preds_df):
        """One bank of the distinct (gene, tissue) forests of the frame, one launch (ad_risk.py:157-176)."""
        keys, ids = {}, []
        for g, t in zip(preds_df.gene_id, preds_df.tissue_id):
            if (g, t) not in keys:
                f = self.ad_preds.get_file_path(g, t) if self.ad_preds else None
                if f is None:
                    raise FileNotFoundError(f"AD predictor not found for gene {g} and tissue {t}")
                keys[(g, t)] = (len(keys), f)
            ids.append(keys[(g, t)][0])
        bank = ForestBank([load_forest(f) for _, f in sorted(keys.values())])
        preds_df["ad_risk"] = bank.predict_proba(np.stack(preds_df.embedding.to_list()), ids)
        if "gene_name" in self.genes_map.columns:
            preds_df["gene_name"] = preds_df["gene_id"].map(self.genes_map["gene_name"])
        return preds_df
