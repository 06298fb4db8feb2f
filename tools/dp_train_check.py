"""Data-parallel training step on real GPUs over NCCL (config #5's collective): each rank runs a small
resolution-preserving encoder/decoder around rpst AdaIN forward+backward on its shard, gradients are averaged
with ONE all-reduce over the flat bucket (rpst.dist.GradBucket); the result must equal the single-process
gradient of the full batch.  torchrun --nproc-per-node N tools/dp_train_check.py"""
import json, os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from rpst.dist import GradBucket, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)


def make_net():
    torch.manual_seed(0)
    return torch.nn.ModuleDict({"enc": torch.nn.Conv2d(3, 32, 3, padding=1), "dec": torch.nn.Conv2d(32, 3, 3, padding=1)}).to(dev)


def loss_of(net, content, style):
    c, s = torch.relu(net["enc"](content)), torch.relu(net["enc"](style))
    out = net["dec"](rpst.adaptive_instance_normalization(c, s))
    return (out - content).square().mean() + rpst.calc_style_loss(torch.relu(net["enc"](out)), s.detach())

g = torch.Generator(device=dev).manual_seed(1)
N = 2 * world
content = torch.rand(N, 3, 256, 512, device=dev, generator=g)
style = torch.rand(N, 3, 256, 512, device=dev, generator=g)
# single-process reference: mean over per-sample-shard losses == mean of rank losses (equal shard sizes)
ref = make_net()
total = 0
for r in range(world):
    lo, hi = shard_range(N, r, world)
    total = total + loss_of(ref, content[lo:hi], style[lo:hi]) / world
total.backward()
net = make_net()
lo, hi = shard_range(N, rank, world)
loss = loss_of(net, content[lo:hi], style[lo:hi])
loss.backward()
bucket = GradBucket(net.parameters())
out = bucket.allreduce_mean({"loss": loss})
err = max(float((p.grad - q.grad).abs().max() / q.grad.abs().max()) for p, q in zip(net.parameters(), ref.parameters()))
# latency of the collective alone on a 3.1 MB bucket (AdaINRPNet decoder)
flat = torch.zeros(784963, device=dev)
for _ in range(5):
    dist.all_reduce(flat)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    dist.all_reduce(flat)
b.record(); torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({"world": world, "grad_rel_err_vs_single_process": err, "loss_mean": float(out["loss"]),
                      "loss_single_process": float(total.detach()), "allreduce_3.1MB_us": a.elapsed_time(b) / 50 * 1e3}), flush=True)
dist.destroy_process_group()
