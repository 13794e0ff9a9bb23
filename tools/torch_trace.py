"""Which ATen kernels does one training step launch, and from where (torch.profiler with stacks)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import synth
from fissure_segmentation_b200.ddp import FlatAdam, FlatDataParallel
from torch.profiler import profile, ProfilerActivity
dev = 'cuda'
torch.manual_seed(0)
import os
CFG = os.environ.get("TRACE_CFG", "train")
if CFG == "configC":
    Bt, Nt, kt, Ct, dyn = 8, 8192, 40, 9, False
elif CFG == "static40":
    Bt, Nt, kt, Ct, dyn = 32, 2048, 40, 3, False
else:
    Bt, Nt, kt, Ct, dyn = 32, 2048, 20, 3, True
model = fs.DGCNNSeg(k=kt, in_features=Ct, num_classes=4, dynamic=dyn).to(dev); model.precision = "bf16"; model.train()
dp = FlatDataParallel(model, n_buckets=2); opt = FlatAdam(dp, lr=1e-3, weight_decay=1e-5)
x, y = synth.make_batch(Bt, Nt, seed=1234, n_features=Ct - 3); x, y = x.to(dev), y.to(dev)
def step():
    dp.zero_grad(); loss = F.cross_entropy(dp(x), y); loss.backward(); dp.finish_backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and (e.device_time_total or getattr(e, "cuda_time_total", 0)) > 0:
        st = [s for s in (e.stack or []) if "fissure_segmentation_b200" in s or "bench" in s or "ddp" in s or "torch_trace" in s]
        rows.append((getattr(e, "self_device_time_total", 0), e.name, str(e.input_shapes)[:60], st[0][-70:] if st else "-"))
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for t, n, sh, st in rows:
    if t <= 0: continue
    agg[(n, sh, st)][0] += 1; agg[(n, sh, st)][1] += t
for (n, sh, st), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%7.1f us x%2d  %-28s %-60s %s" % (t, c, n, sh, st))
print("---- kernels")
kagg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        kagg[e.name[:90]][0] += 1; kagg[e.name[:90]][1] += e.device_time_total
tot = sum(v[1] for v in kagg.values())
print("total kernel us", tot, "kernels", sum(v[0] for v in kagg.values()))
for n, (c, t) in sorted(kagg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[1]) if len(sys.argv) > 1 else 60]:
    print("%7.1f us x%2d  %s" % (t, c, n))
