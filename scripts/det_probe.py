import sys; sys.path.insert(0,'/root/repo')
import torch
from interpolate_antialiasing_b200 import capi
g=torch.Generator().manual_seed(23)
cuda=torch.device('cuda',0)
im=torch.randint(0,256,(90,200,3),dtype=torch.uint8,generator=g).to(cuda)
x=im.permute(2,0,1)[None]
print(x.shape, x.stride(), x.data_ptr()%16, x.dtype)
a=capi.resize_forward(x,(48,56),"linear")
for fl,name in ((capi.FLAG_FORCE_STREAM,'stream'),(capi.FLAG_FORCE_GENERAL,'general'),(capi.FLAG_VMMA,'vmma')):
    try:
        c=capi.resize_forward(x,(48,56),"linear",False,fl); print(name, (a-c).abs().max().item())
    except Exception as e: print(name, 'ERR', e)
out=torch.empty((4,3,48,56),device=cuda).contiguous(memory_format=torch.channels_last)
for i in range(4):
    capi.resize_forward(x,(48,56),"linear",out=out[i:i+1])
torch.cuda.synchronize()
for i in range(4): print(i, torch.equal(out[i:i+1],a), (out[i:i+1]-a).abs().max().item(), out[i:i+1].data_ptr()%16)
