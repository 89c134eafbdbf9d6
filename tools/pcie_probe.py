import torch, time
n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2): run(True)
print("H2D alone GB/s", n / run(False) / 1e9)
t = run(True); print("H2D + D2H concurrently: each GB/s", n / t / 1e9, "ms", t * 1e3)
t = run(True, 16); print("16 chunks: each GB/s", n / t / 1e9, "ms", t * 1e3)
