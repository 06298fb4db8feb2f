"""Jacobi eigensolver time vs batch size (256x256 fp64, content-covariance-like matrices).  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
L = rpst._lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
for batch in (1, 2, 4, 8, 12, 16, 18, 24, 32):
    x = torch.randn(batch, n, 4 * n, device="cuda", dtype=torch.float64, generator=g)
    a = (x @ x.transpose(1, 2) / (4 * n - 1) + torch.eye(n, dtype=torch.float64, device="cuda")).contiguous()
    out = torch.empty_like(a)
    sw = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ws = torch.empty(L.rpst_sym_eig_fn_workspace_bytes(batch, n), dtype=torch.uint8, device="cuda")
    call = lambda: rpst._lib.check(L.rpst_sym_eig_fn(a.data_ptr(), batch, n, 1e-4, out.data_ptr(), None, None, sw.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"n": n, "batch": batch, "ms": e0.elapsed_time(e1), "sweeps_max": int(sw.max())}), flush=True)
