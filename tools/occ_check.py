import sys, torch
sys.path.insert(0, ".")
from variantformer_b200 import ops
DEV="cuda"
H,hd=32,48; d=H*hd
lens_q=[12663]*8; lens_k=[1024]*8
q=torch.randn(sum(lens_q),d,device=DEV).bfloat16(); k=torch.randn(sum(lens_k),d,device=DEV).bfloat16(); v=torch.randn(sum(lens_k),d,device=DEV).bfloat16()
slots=ops.SlotMap(lens_q,DEV,k_lens=lens_k)
o=ops.attention_mc(q,k,v,slots,H,hd,None); torch.cuda.synchronize(); print("ok")
