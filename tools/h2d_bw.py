"""PCIe copy bandwidth of the box with pinned memory (context for bench.py's e2e line)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 8, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, both in (('h2d alone', False), ('h2d with concurrent d2h', True)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{name}: {4 * n / dt / 1e9:.1f} GB/s')
for chunk in (64 << 20, 190 << 20, 512 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for o in range(0, n, chunk):
        d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'h2d in {chunk >> 20} MB chunks: {n / dt / 1e9:.1f} GB/s')
