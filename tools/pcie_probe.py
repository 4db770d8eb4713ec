#!/usr/bin/env python3
"""PCIe copy rates of the box for the sizes of the end-to-end leg (pinned host memory, CUDA events)."""
import torch, json
dev = torch.device('cuda:0')
res = {}
for name, nbytes in (('d2h_29MB', 29427712), ('h2d_15MB', 14713856), ('d2h_3.7MB', 29427712 // 8)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src, dst = (d, h) if name.startswith('d2h') else (h, d)
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res[name] = {'ms': ms, 'GB/s': nbytes / ms / 1e6}
# both directions at once
h1 = torch.empty(29427712, dtype=torch.uint8).pin_memory(); d1 = torch.empty(29427712, dtype=torch.uint8, device=dev)
h2 = torch.empty(14713856, dtype=torch.uint8).pin_memory(); d2 = torch.empty(14713856, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
res['both_directions_ms_per_pair'] = (time.perf_counter() - t0) * 1e3 / 20
print(json.dumps(res, indent=1))
