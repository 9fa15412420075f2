import torch, time
x = torch.randn(2, 65536, 1000, dtype=torch.float64, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/n
b = x.numel()*8
print("sum   %.1f GB/s" % (b/t(lambda: x.sum())/1e6))
print("amax(dim=-1) %.1f GB/s" % (b/t(lambda: x.amax(dim=-1))/1e6))
print("max   %.1f GB/s" % (b/t(lambda: x.max())/1e6))
y = torch.empty_like(x)
print("copy  %.1f GB/s (r+w)" % (2*b/t(lambda: y.copy_(x))/1e6))
