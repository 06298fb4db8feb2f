"""Cycle counters of the flash attention kernel's MMA thread and one softmax thread (pair 0)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst

dev = torch.device("cuda")
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device=dev).manual_seed(4)
f = torch.randn(b, 512, side, side, device=dev, generator=g) * 0.3
k = torch.randn(b, 512, side, side, device=dev, generator=g) * 0.3
v = torch.randn(b, 512, side, side, device=dev, generator=g)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
names_mma = ["total", "wait_full", "wait_peer_full", "wait_s_empty", "wait_p_full", "wait_q", "tiles"]
names_sm = ["total", "wait_s_full", "pass1", "max_exchange", "wait_o_done", "rescale", "pass2", "epilogue"]
for prec in ("bf16", "fp32"):
    rpst.attention_core(f, k, v, precision=prec)
    torch.cuda.synchronize()
    prof.zero_()
    rpst.set_tuning("attn_flash_prof", prof.data_ptr())
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rpst.attention_core(f, k, v, precision=prec)
    e.record()
    torch.cuda.synchronize()
    rpst.set_tuning("attn_flash_prof", 0)
    c = prof.cpu().tolist()
    tiles = max(c[6], 1)
    print(json.dumps({"prec": prec, "L": side * side, "b": b, "ms_incl_pack": a.elapsed_time(e),
                      "mma_thread_cycles_per_tile": {n: round(c[i] / tiles, 1) for i, n in enumerate(names_mma[:-1])},
                      "softmax_thread_cycles_per_tile": {n: round(c[8 + i] / tiles, 1) for i, n in enumerate(names_sm)},
                      "tiles": c[6]}), flush=True)
