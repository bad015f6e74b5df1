import torch, time
for mb in (32, 64):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): d.copy_(h, non_blocking=True)
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {mb} MiB pinned: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
