import torch, sys
sys.path.insert(0,'/root/repo')
dev=torch.device('cuda',0)
x=torch.rand((256,3,1080,1920),device=dev)
def t(fn,iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/iters
nbytes=x.numel()*4
ms=t(lambda: x.sum()); print(f"torch.sum        {ms:.3f} ms {nbytes/ms/1e6:.0f} GB/s")
ms=t(lambda: x.amax()); print(f"torch.amax       {ms:.3f} ms {nbytes/ms/1e6:.0f} GB/s")
y=torch.empty_like(x)
ms=t(lambda: y.copy_(x)); print(f"copy (r+w)       {ms:.3f} ms {2*nbytes/ms/1e6:.0f} GB/s")
ms=t(lambda: y.zero_()); print(f"memset (w)       {ms:.3f} ms {nbytes/ms/1e6:.0f} GB/s")
xs=x.view(-1)
ms=t(lambda: torch.dot(xs[:2**31-8], xs[:2**31-8]) if False else xs.norm()); print(f"torch.norm       {ms:.3f} ms {nbytes/ms/1e6:.0f} GB/s")
