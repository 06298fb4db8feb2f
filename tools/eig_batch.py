"""16 x 256x256 batched matrix square roots (the Jacobi solve of BASELINE config #3) — profiler target."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
n, batch = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device="cuda").manual_seed(n)
a = torch.randn(batch, n, 2 * n, device="cuda", dtype=torch.float64, generator=g)
a = a @ a.transpose(1, 2) / (2 * n)
for _ in range(2):
    r = rpst.matrix_sqrt(a)
torch.cuda.synchronize()
print("ok", float(r[0, 0, 0]))
