import sys, os
sys.path.insert(0, "/root/repo")
import torch
from gaot_b200 import _lib
from gaot_b200._lib import ptr
L = _lib.lib()
for (B,T,H,Hkv,hd) in [(2,256,2,2,32),(1,1024,8,8,32),(2,200,4,2,32),(1,384,2,2,64),(1,130,2,1,16)]:
    W=(H+2*Hkv)*hd
    qkv=torch.randn(B*T,W,device="cuda").to(torch.bfloat16)
    o=torch.empty(B*T,H*hd,dtype=torch.bfloat16,device="cuda"); lse=torch.empty(B,H,T,device="cuda")
    st=torch.cuda.current_stream().cuda_stream
    rc=L.gaot_attention_fwd_bf16(ptr(qkv),ptr(o),ptr(lse),B,T,H,Hkv,hd,W,1.0/hd**0.5,st)
    torch.cuda.synchronize(); print("fwd rc",rc)
    do=torch.randn_like(o); dqkv=torch.empty_like(qkv); delta=torch.empty_like(lse); dq=torch.empty(B*T,H*hd,device="cuda")
    rc=L.gaot_attention_bwd_bf16(ptr(qkv),ptr(o),ptr(lse),ptr(do),ptr(delta),ptr(dq),ptr(dqkv),B,T,H,Hkv,hd,W,1.0/hd**0.5,st)
    print("bwd rc",rc, (B,T,H,Hkv,hd)); 
    try:
        torch.cuda.synchronize(); print("sync ok", float(dqkv.float().abs().mean()))
    except Exception as e:
        print("sync err", e); break
