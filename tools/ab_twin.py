"""A/B of twin-chunk apply items (plain AdaIN / mean_variance_norm on the TMA kernel), same process.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
dev = torch.device("cuda")
H = 512
for batch, ch in ((16, 256), (32, 64)):
    c = torch.relu(torch.randn(batch, ch, H, H, device=dev) + 0.5)
    s = torch.relu(torch.randn(batch, ch, H, H, device=dev) * 2 + 1)
    E = c.numel() * 4
    for rep in range(2):
        for mode in (0, 1):
            rpst.set_tuning("adain_twin_apply", mode)
            for name, fn, nb in (("adain", lambda: rpst.adaptive_instance_normalization(c, s), 3), ("mvn", lambda: rpst.mean_variance_norm(c), 2)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 10
                print(json.dumps({"op": name, "batch": batch, "ch": ch, "twin": mode, "rep": rep, "ms": ms, "GBs": nb * E / ms / 1e6}), flush=True)
    del c, s
