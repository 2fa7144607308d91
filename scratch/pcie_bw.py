"""PCIe ceiling of the box: pinned H2D, D2H, and both at once (torch, two streams)."""
import torch, time
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
run(True, True, 2)
print("H2D alone  %.1f GB/s" % run(True, False))
print("D2H alone  %.1f GB/s" % run(False, True))
print("both       %.1f GB/s per direction" % run(True, True))
