import torch, time
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t) / 1e9
print("H2D pinned GB/s", bw(lambda: d.copy_(h, non_blocking=True)))
print("D2H pinned GB/s", bw(lambda: h.copy_(d, non_blocking=True)))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidirectional GB/s (each dir)", bw(both))
p = torch.empty(n, dtype=torch.uint8)
print("H2D pageable GB/s", bw(lambda: d.copy_(p), 3))
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -12; lscpu | grep -E 'Model name|NUMA|Socket'", shell=True, capture_output=True, text=True).stdout)
