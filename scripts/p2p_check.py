"""torchrun --nproc-per-node N scripts/p2p_check.py: the one-shot peer-memory all-reduce against NCCL on random vectors,
then its latency next to NCCL's for the packed-buffer size of c3/c5 (1025 floats)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from bijax_b200.parallel import P2PAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 1025
op = P2PAllReduce(n, None, dev)
ok = True
for it in range(200):
    g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
    v = torch.randn(n, device=dev, generator=g)
    a = v.clone(); op(a)
    b = v.clone(); dist.all_reduce(b)
    # reference: gather everything and add in rank order
    parts = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(parts, v)
    want = torch.zeros_like(v)
    for p in parts: want = want + p
    ok = ok and torch.equal(a, want) and torch.allclose(a, b, rtol=1e-6, atol=1e-6)
def lat(fn, iters=2000):
    x = torch.randn(n, device=dev)
    for _ in range(50): fn(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn(x)
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters
t_p2p = lat(op)
t_nccl = lat(lambda x: dist.all_reduce(x))
res = torch.tensor([float(ok), t_p2p, t_nccl], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN if False else dist.ReduceOp.SUM)
if rank == 0:
    print(f"world={world} exact_vs_rank_order_sum_and_close_to_nccl={res[0].item() == world}  p2p {res[1].item()/world:.2f} us  nccl {res[2].item()/world:.2f} us per all-reduce of {n} floats (back to back)")
op.close()
dist.destroy_process_group()
