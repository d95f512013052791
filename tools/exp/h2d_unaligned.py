import torch, time
n = 46_000_000
host = torch.empty(n + 4096, dtype=torch.uint8).pin_memory()
dev = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
for off_s, off_d in ((0, 0), (1, 0), (3, 0), (64, 0), (77, 0), (1, 1), (4, 0), (16, 0)):
    for _ in range(2):
        dev[off_d:off_d + n].copy_(host[off_s:off_s + n], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dev[off_d:off_d + n].copy_(host[off_s:off_s + n], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(off_s, off_d, f"{n * 5 / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
