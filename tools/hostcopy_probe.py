import time, torch, numpy as np
torch.cuda.init(); x = torch.randn(5, 4096, 2313, dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
mb = x.numel()*8/1e6
def T(f, n=3):
    ts=[]
    for _ in range(n):
        torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
    return min(ts)*1e3
print("MB", mb)
print("cpu() pageable ms", T(lambda: x.cpu()))
print("pinned alloc ms", T(lambda: torch.empty(x.shape, dtype=torch.float64, pin_memory=True)))
p = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
print("D2H pinned ms", T(lambda: p.copy_(x, non_blocking=True)))
print("np.empty+copy from pinned ms", T(lambda: np.array(p.numpy(), copy=True)))
out = np.empty(tuple(x.shape))
print("copy into existing (touched) ms", T(lambda: np.copyto(out, p.numpy())))
print("H2D pageable 75MB ms", T(lambda: torch.as_tensor(out[0]).cuda()))
print("H2D pinned 75MB ms", T(lambda: p[0].cuda()))
import os; print("cpus", os.cpu_count())
