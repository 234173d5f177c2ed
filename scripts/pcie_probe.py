"""PCIe ceiling of the box: pinned H2D, D2H and both at once (two streams), sizes of one HLS tile."""
import time, torch
n_in, n_out = 256_000_000, 67_000_000
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); di = torch.empty(n_in, dtype=torch.uint8, device='cuda')
ho = torch.empty(n_out, dtype=torch.uint8).pin_memory(); do = torch.empty(n_out, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for _ in range(2): run(True, True)
t = run(True, False); print(f'H2D 256 MB: {t*1e3:.2f} ms  {n_in/t/1e9:.1f} GB/s')
t = run(False, True); print(f'D2H  67 MB: {t*1e3:.2f} ms  {n_out/t/1e9:.1f} GB/s')
t = run(True, True); print(f'both       : {t*1e3:.2f} ms  {(n_in+n_out)/t/1e9:.1f} GB/s total')
# chunked H2D (10 copies per strip, 4 strips) to see the per-copy overhead
chunks = list(hi.chunk(40)); dch = list(di.chunk(40))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    for a, b in zip(chunks, dch): b.copy_(a, non_blocking=True)
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
print(f'H2D in 40 chunks: {t*1e3:.2f} ms  {n_in/t/1e9:.1f} GB/s')
