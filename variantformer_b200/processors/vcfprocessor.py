"""VCFProcessor — the vcf2exp API with the reference's method names and return shapes
(processors/vcfprocessor.py:24-61, :217-277), plus a Lightning-free `Trainer` exposing `.predict` / `.precision`."""
import logging
import os

import pandas as pd
import torch
from torch.utils.data import DataLoader

from ..datasets.vcfdataset import LocalGeneManifest, VCFDataset, collate_fn_batching
from ..utils.config import CONFIG_DIR, PACKAGE_ROOT, VOCAB_DIR, load_yaml
from .model_manager import ModelManager

log = logging.getLogger(__name__)


class Trainer:
    """Stand-in for lightning.pytorch.Trainer(accelerator='gpu', devices=1, precision=...): single process, current
    CUDA stream, `predict(model, dataloader, ckpt_path=None) -> list[dict]`."""

    def __init__(self, accelerator="gpu", devices=1, logger=False, precision="bf16-mixed", enable_checkpointing=False,
                 **_):
        assert devices == 1, "one process drives one GPU; use variantformer_b200.parallel for multi-GPU sharding"
        self.precision = precision

    def predict(self, model, dataloaders, ckpt_path=None):
        if ckpt_path and os.path.exists(ckpt_path):             # Lightning reloads the weights (vcfprocessor.py:262)
            ck = torch.load(ckpt_path, map_location="cpu")
            model.load_state_dict(ck["state_dict"] if "state_dict" in ck else ck)
            model.to("cuda")
        model.trainer = self
        model.eval()
        return [model.predict_step(batch, i) for i, batch in enumerate(dataloaders)]


class VCFProcessor:
    def __init__(self, model_class: str = "v4_pcg", base_dir=None, gene_cre_manifest=None, model_overrides=None):
        """base_dir: where relative artifact paths (`_artifacts/...`) are resolved (reference: the repo root).
        gene_cre_manifest: object with get_file_path(gene_id); default = <base_dir>/_artifacts/gene_cre_manifests/."""
        base_dir = base_dir or PACKAGE_ROOT
        self.config_location = CONFIG_DIR
        self.model_config = load_yaml(os.path.join(CONFIG_DIR, "vf_model.yaml"))[model_class]
        if model_overrides:
            self.model_config.model.update(model_overrides)
        self.tissue_vocab = load_yaml(os.path.join(VOCAB_DIR, "tissue_vocab.yaml"))
        self.vcf_loader_config = load_yaml(os.path.join(CONFIG_DIR, "vcfloader.yaml"))
        self.gene_cre_manifest = gene_cre_manifest or LocalGeneManifest(
            os.path.join(base_dir, "_artifacts", "gene_cre_manifests"))

        def fix(node, key):
            if node.get(key) and not os.path.isabs(node[key]):
                node[key] = os.path.join(base_dir, node[key])
        fix(self.vcf_loader_config, "CRE_BED"); fix(self.vcf_loader_config, "fasta_path")
        fix(self.model_config.dataset, "gencode_v24"); fix(self.model_config.model, "checkpoint_path")
        fix(self.model_config.model.cre_tokenizer, "path"); fix(self.model_config.model.gene_tokenizer, "path")
        assert torch.cuda.is_available(), "GPU is not available"
        self.accelerator = "gpu"

    def get_tissues(self):
        return self.tissue_vocab.keys()

    def get_genes(self):
        return pd.read_csv(self.model_config.dataset.gencode_v24)

    def create_data(self, vcf_path: str, query_df: pd.DataFrame, **kwargs):
        cfg = dict(self.vcf_loader_config.dataloader)
        cfg.update(kwargs)
        # items are produced by CUDA kernels in the main process: no worker processes, no pinning of device tensors
        cfg.update(num_workers=0, pin_memory=False); cfg.pop("prefetch_factor", None)
        ds = self.model_config.dataset
        dataset = VCFDataset(max_length=ds.max_length, max_chunks=ds.max_chunks, cre_neighbour_hood=ds.cre_neighbour_hood,
                             gencode_v24=ds.gencode_v24, gene_cre_manifest=self.gene_cre_manifest,
                             gene_upstream_neighbour_hood=ds.gene_upstream_neighbour_hood,
                             gene_downstream_neighbour_hood=ds.gene_downstream_neighbour_hood, query_df=query_df,
                             fasta_path=self.vcf_loader_config.fasta_path, vcf_path=vcf_path)
        return dataset, DataLoader(dataset, collate_fn=collate_fn_batching, **cfg)

    def load_model(self):
        model, checkpoint_path = ModelManager(self.model_config.model).load_model()
        trainer = Trainer(accelerator=self.accelerator, devices=1, logger=False,
                          precision=self.model_config.model.precision, enable_checkpointing=False)
        return model, checkpoint_path, trainer

    def predict(self, model, checkpoint_path, trainer, dataloader, vcf_dataset):
        predictions = trainer.predict(model, dataloader, ckpt_path=checkpoint_path)
        return self.format_output(vcf_dataset.query_df, predictions)

    def format_output(self, df, predictions):
        pred_exp, embd = [], []
        for p in predictions:
            pred_exp.extend(p["pred_gene_exp"]); embd.extend(p["embeddings"])
        assert len(df) == len(pred_exp), "DataFrame and predictions length mismatch"
        df["predicted_expression"] = pd.Series(pred_exp, index=df.index)
        df["embeddings"] = pd.Series(embd, index=df.index)
        return df
