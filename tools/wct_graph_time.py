"""wct_fuse at config #3: eager launches vs one CUDA-graph replay.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
c, s = R.synth_features((16, 256, 512, 512), cfg=3, device="cuda")
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
eager = t(lambda: rpst.wct_fuse(c, s))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = rpst.wct_fuse(c, s)
graph = t(lambda: g.replay())
print(json.dumps({"op": "wct_fuse 16x256x512x512", "eager_ms_per_sample": eager / 16, "graph_ms_per_sample": graph / 16}))
