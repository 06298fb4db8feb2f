import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
L = rpst._lib.lib()
c, s = R.synth_features((4, 256, 256, 256), cfg=3, device="cuda")
x = c.reshape(4, 256, -1).double()
x = x - x.mean(2, keepdim=True)
cov = x @ x.transpose(1, 2) / (x.shape[2] - 1) + torch.eye(256, dtype=torch.float64, device="cuda")
xs = s.reshape(4, 256, -1).double(); xs = xs - xs.mean(2, keepdim=True)
covs = (xs @ xs.transpose(1, 2) / (xs.shape[2] - 1)).contiguous()
for name, a in (("content cov + I", cov.contiguous()), ("style cov", covs)):
    sw = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = torch.empty_like(a)
    ws = torch.empty(L.rpst_sym_eig_fn_workspace_bytes(4, 256), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        rpst._lib.check(L.rpst_sym_eig_fn(a.data_ptr(), 4, 256, 1e-4, out.data_ptr(), None, None, sw.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rpst._lib.check(L.rpst_sym_eig_fn(a.data_ptr(), 4, 256, 1e-4, out.data_ptr(), None, None, sw.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
    print(name, "sweeps", sw.tolist(), "ms", e0.elapsed_time(e1), "err", R.rel_l2(out[0], R.matrix_sqrt(a[0].cpu())))
