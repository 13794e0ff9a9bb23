"""torchrun --nproc-per-node 2 tools/check_fused_tail.py: the peer-memory tail (sum over ranks + Adam in one kernel,
fs_adam_step_peers) against NCCL: the reduced gradient must equal all_reduce(sum) of the local tail gradients bit for bit
(same order of summation for 2 ranks), the updated parameters must equal a reference Adam step on that sum, and all ranks
must hold identical parameters."""
import os, sys, torch, torch.distributed as dist, torch.nn.functional as F
sys.path.insert(0, '.')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import synth
from fissure_segmentation_b200.ddp import FlatAdam, FlatDataParallel
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
world = dist.get_world_size()
torch.manual_seed(0)
m = fs.DGCNNSeg(k=20, in_features=3, num_classes=4).to(dev); m.precision = "bf16"; m.train()
MODE = os.environ.get("FUSED", "1")          # 1: EdgeConv bucket from peer memory; all: every gradient
dp = FlatDataParallel(m, n_buckets=2, tail_share=0.04, fused_tail="all" if MODE == "all" else True)
assert dp.symm is not None, "symmetric memory not available"
opt = FlatAdam(dp, lr=1e-3, weight_decay=1e-5)
s, e, _ = dp.buckets[-1]
opt.tail_sum = torch.zeros(e - s, device=dev)
x, y = synth.make_batch(4, 1024, seed=100 + rank); x, y = x.to(dev), y.to(dev)
for step in range(3):
    dp.zero_grad(); F.cross_entropy(dp(x), y).backward(); dp.finish_backward()
    expect = dp.flat_grad[s:e].clone()
    dist.all_reduce(expect)                              # NCCL sum of the local tail gradients
    p0, m0, v0 = dp.flat_param[s:e].clone(), opt.exp_avg[s:e].clone(), opt.exp_avg_sq[s:e].clone()
    head_expect = dp.flat_grad[:s].clone()               # already reduced by the bucket all-reduce
    opt.step()
    torch.cuda.synchronize()
    dsum = float((opt.tail_sum - expect).abs().max())
    # reference Adam on the NCCL sum (torch.optim.Adam semantics with L2 weight decay)
    t = step + 1
    g = expect / world + 1e-5 * p0
    m1 = m0 + (1 - 0.9) * (g - m0); v1 = 0.999 * v0 + (1 - 0.999) * g * g
    p1 = p0 - (1e-3 / (1 - 0.9 ** t)) * (m1 / (v1.sqrt() / (1 - 0.999 ** t) ** 0.5 + 1e-8))
    dpar = float((dp.flat_param[s:e] - p1).abs().max())
    if rank == 0:
        print("step %d: max |peer sum - nccl sum| = %.3e (scale %.3e), max |param - reference Adam| = %.3e"
              % (step, dsum, float(expect.abs().max()), dpar))
    assert dsum <= 1e-6 * float(expect.abs().max()) + 1e-12, dsum
    assert dpar <= 1e-6, dpar
others = [torch.empty_like(dp.flat_param) for _ in range(world)]
dist.all_gather(others, dp.flat_param)
assert all(torch.equal(o, others[0]) for o in others), "ranks diverged"
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("fused tail (%s) ok: ranks hold identical parameters; tail = %d of %d parameters" % (MODE, e - s, dp.flat_param.numel()))
