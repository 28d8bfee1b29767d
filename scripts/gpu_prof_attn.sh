#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/attn_prof.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from svd_xtend_b200 import raw
bf16=torch.bfloat16
nseq,S,heads=14,2560,5
C=heads*64
qkv=torch.randn(nseq*S,3*C,device='cuda').to(bf16)
o=torch.empty(nseq*S,C,device='cuda',dtype=bf16); lse=torch.empty(nseq*S,heads,device='cuda')
dout=torch.randn(nseq*S,C,device='cuda').to(bf16); dqkv=torch.empty_like(qkv); delta=torch.empty_like(lse)
for _ in range(3):
    raw.attention_fwd(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],o,heads=heads,S=S,nseq=nseq,lse=lse)
    raw.attention_bwd(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],o,dout,dqkv[:,:C],dqkv[:,C:2*C],dqkv[:,2*C:],lse,delta,heads=heads,S=S,nseq=nseq)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    raw.attention_fwd(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],o,heads=heads,S=S,nseq=nseq,lse=lse)
e1.record(); torch.cuda.synchronize(); print("fwd us", e0.elapsed_time(e1)*100)
e0.record()
for _ in range(10):
    raw.attention_bwd(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],o,dout,dqkv[:,:C],dqkv[:,C:2*C],dqkv[:,2*C:],lse,delta,heads=heads,S=S,nseq=nseq)
e1.record(); torch.cuda.synchronize(); print("bwd us", e0.elapsed_time(e1)*100)
PY
python /tmp/attn_prof.py > gpurun_out/attn_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/prof_attn_fwd_r1 python /tmp/attn_prof.py >> gpurun_out/attn_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq -s 3 -c 1 -f -o gpurun_out/prof_attn_dq_r1 python /tmp/attn_prof.py >> gpurun_out/attn_prof.log 2>&1
tail -5 gpurun_out/attn_prof.log
