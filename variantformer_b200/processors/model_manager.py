"""ModelManager — builds the two seq2reg encoders + the seq2gene model and loads checkpoints, with the reference's
contract (processors/model_manager.py:24-27, :62-121): `ModelManager(config).load_model() -> (model, ckpt_path)`;
raises ValueError("Checkpoint not found"); model is eval(), on CUDA, `model.vep = False`."""
import logging
import os

import torch

from ..seq2gene.model_combined_modulator import Seq2GenePredictorCombinedModulator
from ..seq2reg.model import Seq2RegPredictor
from ..utils import random_init
from ..utils.config import Config

log = logging.getLogger(__name__)


class ModelManager:
    """config = vf_model.yaml[<model_class>].model (Config, dict or OmegaConf node).
    Extra key `random_init: <seed>` (not in the reference) skips every torch.load and fills the reference-shaped
    state_dict with seeded random weights — the offline benchmarking path; the default behaviour is untouched."""

    def __init__(self, config):
        self.config = config if hasattr(config, "copy") and hasattr(config, "get") else Config(dict(config))
        self.model = None
        if not torch.cuda.is_available():
            raise RuntimeError("variantformer_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.device = "cuda"

    @staticmethod
    def _load_seq2reg(path) -> Seq2RegPredictor:
        chk = torch.load(path, map_location="cpu", weights_only=False)
        m = Seq2RegPredictor(**chk["hyper_parameters"])
        m.load_state_dict(chk["state_dict"])
        return m

    def load_model(self):
        cfg = self.config.copy()
        seed = cfg.pop("random_init", None) if isinstance(cfg, dict) else None
        hp_override = cfg.pop("seq2reg_hyper_parameters", None) if isinstance(cfg, dict) else None
        if seed is None:
            log.info("Loading Seq2Reg model...")
            seq2reg = self._load_seq2reg(cfg["cre_tokenizer"]["path"])
            log.info("Loading Seq2Reg gene model...")
            seq2reg_gene = self._load_seq2reg(cfg["gene_tokenizer"]["path"])
        else:
            hp = dict(hp_override or random_init.SEQ2REG_HP)
            seq2reg, seq2reg_gene = Seq2RegPredictor(**hp), Seq2RegPredictor(**hp)
        for k in ("cre_tokenizer", "gene_tokenizer"):
            if k in cfg:
                del cfg[k]
        cfg["token_dim"] = seq2reg.hparams.embedding_dim
        classes = {"Seq2GenePredictorCombinedModulator": Seq2GenePredictorCombinedModulator}
        name = cfg.get("model_class", "Seq2GenePredictor")
        if name not in classes:
            raise NotImplementedError(f"model_class {name!r}: only Seq2GenePredictorCombinedModulator (the class both "
                                      "vf_model.yaml entries select) is implemented on the B200 path")
        checkpoint_path = self.config.get("checkpoint_path")
        log.info("Creating Seq2Gene model...")
        model = classes[name](cre_tokenizer=seq2reg, gene_tokenizer=seq2reg_gene, **cfg)
        log.info(f"Total number of parameters: {sum(p.numel() for p in model.parameters()):,}")
        if seed is None:
            if not checkpoint_path or not os.path.exists(checkpoint_path):
                raise ValueError("Checkpoint not found")
            log.info(f"Loading checkpoint from {checkpoint_path}")
            ck = torch.load(checkpoint_path, map_location="cpu")
            model.load_state_dict(ck["state_dict"] if "state_dict" in ck else ck)
        else:
            model.load_state_dict(random_init.make_state_dict(dict(cfg), dict(seq2reg.hparams), seed=int(seed)))
        model.eval()
        model.to(self.device)
        model.vep = False
        self.model = model
        log.info(f"Model loaded successfully on {self.device}")
        return model, checkpoint_path
