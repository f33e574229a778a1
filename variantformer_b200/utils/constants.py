"""Domain constants of the hot path (values follow the reference's utils/constants.py:2-31,95-100)."""
IUPAC_ALPHABET = "ACGTRYSWKMBDHV"          # the 14 codes BPEEncoder.normalize keeps; 'N' is NOT one of them
IUPAC_CODES = {c: None for c in IUPAC_ALPHABET}
REF_CREs = [
    "CTCF-only,CTCF-bound", "DNase-H3K4me3", "DNase-H3K4me3,CTCF-bound", "PLS", "PLS,CTCF-bound",
    "dELS", "dELS,CTCF-bound", "pELS", "pELS,CTCF-bound",
]
MAP_REF_CRE_TO_IDX = {cre: idx for idx, cre in enumerate(REF_CREs)}
SPECIAL_TOKENS = {"pad_token": "<pad>", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>"}
PAD_TOKEN_ID = 0
