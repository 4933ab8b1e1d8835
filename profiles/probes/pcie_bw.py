"""Host<->device copy bandwidth with pinned memory (what bounds the e2e number's H2D / D2H legs)."""
import torch, time
n = 400 * 2**20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h.copy_(d, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D {n/a/1e9:.1f} GB/s, D2H {n/b/1e9:.1f} GB/s, both at once: {n/c/1e9:.1f} GB/s each direction")
