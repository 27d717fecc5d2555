"""PCIe copy bandwidth of this box (pinned host memory), for reading the e2e number."""
import torch
n = 48 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, k=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


print("H2D 48 MiB: %.1f GB/s" % (n / t(lambda: d.copy_(h, non_blocking=True)) / 1e6))
print("D2H 48 MiB: %.1f GB/s" % (n / t(lambda: h.copy_(d, non_blocking=True)) / 1e6))


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


ms = t(both)
print("H2D + D2H concurrently, 48 MiB each: %.2f ms -> %.1f GB/s per direction" % (ms, n / ms / 1e6))
