import sys
sys.path.insert(0, "/root/repo")
import torch
from oracle import clip_search_oracle as orc
from summer_clip_b200 import ops
nq=777
banks = orc.synthetic_banks(nq, 5000, 256, 300, seed=5, sigma=0.5, sigma_text=0.8, shared=3.0)
Q, K, L, T = (banks[n].cuda() for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
lo, hi = 389, 777
zf = ops.zero_shot_logits(Q, True, T)
zs = ops.zero_shot_logits(Q[:, lo:hi], True, T)
zsimt = ops.zero_shot_logits(Q[:, lo:hi], True, T, tensor_cores=False)
print("zero-shot slice TC vs full", (zs - zf[lo:hi]).abs().max().item(), "SIMT vs full", (zsimt - zf[lo:hi]).abs().max().item())
qn = ops.normalize_cast(Q, True)
kn = ops.normalize_cast(K, True)
vt = ops.values_prepare(L, 300, softmax_scale=orc.CLIP_SCALE*0.1)
of = ops.attn_fwd(qn, kn, vt, 5000, 300, 5.5)
os_ = ops.attn_fwd(qn[lo:hi], kn, vt, 5000, 300, 5.5)
print("attn slice vs full", (os_ - of[lo:hi]).abs().max().item() / of.abs().max().item())
oc = ops.attn_fwd(qn[lo:hi].clone(), kn, vt, 5000, 300, 5.5)
print("attn slice(clone) vs full", (oc - of[lo:hi]).abs().max().item() / of.abs().max().item())
# key shards
p0 = ops.attn_fwd(qn, kn[:2560].contiguous(), ops.values_prepare(L, 300, idx=torch.arange(0,2560,device='cuda'), softmax_scale=orc.CLIP_SCALE*0.1), 2560, 300, 5.5)
p1 = ops.attn_fwd(qn, kn[2560:].contiguous(), ops.values_prepare(L, 300, idx=torch.arange(2560,5000,device='cuda'), softmax_scale=orc.CLIP_SCALE*0.1), 2440, 300, 5.5)
print("shards vs full rows<389", ((p0+p1)[:389]-of[:389]).abs().max().item()/of.abs().max().item(), "rows>=389", ((p0+p1)[389:]-of[389:]).abs().max().item()/of.abs().max().item())
ep = ops.epilogue(zf[lo:hi].contiguous(), of[lo:hi].contiguous(), [0.5, 2.0], labels=banks["test_labels"].cuda()[lo:hi].contiguous(), want_logits=True)
ef = ops.epilogue(zf, of, [0.5, 2.0], labels=banks["test_labels"].cuda(), want_logits=True)
print("epilogue slice vs full", (ep["logits"] - ef["logits"][:, lo:hi]).abs().max().item())
ep2 = ops.epilogue(zf[lo:hi], of[lo:hi], [0.5, 2.0], labels=banks["test_labels"].cuda()[lo:hi].contiguous(), want_logits=True)
print("epilogue slice(views) vs full", (ep2["logits"] - ef["logits"][:, lo:hi]).abs().max().item())
