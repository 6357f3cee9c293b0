import torch
dev = torch.device("cuda", 0)
for nbytes in (805_306_368, 2_000_000_000):
    x = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
    for name, fn in (("fill_", lambda: x.fill_(float("nan"))), ("zero_", lambda: x.zero_())):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(nbytes, name, round(ms, 4), "ms", round(nbytes / ms / 1e6, 1), "GB/s")
