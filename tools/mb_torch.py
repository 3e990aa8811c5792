import torch, time
DEV="cuda"
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
s=torch.randn(3038817,64,device=DEV); o=torch.randn(1013889,64,device=DEV); out=torch.empty_like(o)
ms=t(lambda: s.sum()); print(f"sum(s) {ms:.4f} ms {s.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: out.copy_(o)); print(f"copy o {ms:.4f} ms {2*o.numel()*4/ms/1e6:.0f} GB/s")
s2=torch.empty_like(s)
ms=t(lambda: s2.copy_(s)); print(f"copy s {ms:.4f} ms {2*s.numel()*4/ms/1e6:.0f} GB/s")
idx=torch.arange(3038817,device=DEV)//3
ms=t(lambda: torch.index_select(o,0,idx)); print(f"index_select rows {ms:.4f} ms read~{(s.numel()*4)/ms/1e6:.0f} GB/s written")
ms=t(lambda: torch.add(s, s2)); print(f"add s+s2 {ms:.4f} ms {3*s.numel()*4/ms/1e6:.0f} GB/s")
