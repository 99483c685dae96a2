import torch, time
a = torch.empty(1<<30, dtype=torch.bfloat16, device='cuda'); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); time.sleep(1.0)
n = 400
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n+1)]
ev[0].record()
for k in range(n):
    b.copy_(a); ev[k+1].record()
torch.cuda.synchronize()
ts = [ev[k].elapsed_time(ev[k+1]) for k in range(n)]
gb = 2*a.numel()*2/1e9
import statistics
print("first 10 GB/s:", [round(gb/t*1e3) for t in ts[:10]])
for lo in (0, 50, 100, 200, 300): print(lo, "..", lo+50, "median GB/s", round(gb/statistics.median(ts[lo:lo+50])*1e3), "cum ms", round(sum(ts[:lo+50])))
