"""In-situ per-kernel device times of one bench step (torch.profiler / CUPTI activity records, which also cover kernels
launched by CUDA-graph replay).  Unlike an ncu launch list the kernels run back to back with warm caches, exactly as in
the timed pass; unlike the CUDA-event stage ranges of bench.py nothing is launched eagerly.
usage: python scripts/kernel_times.py <workload> [images per GPU]  ->  table on stdout"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.cuda.set_device(0)
wl = bench.Workload(name, 0, 1, torch.device('cuda:0'), 'bf16', batch=batch)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); wl.step(); e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wl.step()
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t = ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
        n, tot, mx = agg.get(ev.name, (0, 0.0, 0.0))
        agg[ev.name] = (n + 1, tot + t, max(mx, t))
total = sum(v[1] for v in agg.values())
print(f'# {name}: {wl.n_local} images, step {step_ms:.3f} ms (CUDA events, unprofiled); kernel time under the profiler {total / 1e3:.3f} ms')
print(f'{"kernel":100s} {"count":>6s} {"total ms":>10s} {"avg us":>9s} {"max us":>9s} {"share":>6s}')
for k, (n, tot, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:100]:100s} {n:6d} {tot / 1e3:10.3f} {tot / n:9.2f} {mx:9.2f} {tot / total:6.3f}')

# ---- idle time between consecutive kernels (where the step is not covered by kernel time)
evs = sorted([(ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
              if ev.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e[0])
gaps, by_prev = [], {}
for (s0, e0, n0), (s1, e1, n1) in zip(evs, evs[1:]):
    g = s1 - e0
    if g > 0:
        gaps.append((g, n0, n1))
        c, t = by_prev.get(n0[:60], (0, 0.0))
        by_prev[n0[:60]] = (c + 1, t + g)
print(f'\n# idle time between kernels: {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps; by preceding kernel:')
for k, (c, t) in sorted(by_prev.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'  {t / 1e3:8.3f} ms in {c:5d} gaps after {k}')
print('# largest gaps:')
for g, n0, n1 in sorted(gaps, key=lambda g: -g[0])[:12]:
    print(f'  {g:9.1f} us  {n0[:50]}  ->  {n1[:50]}')
