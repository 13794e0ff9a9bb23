import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import model_fixtures as MF
from test_secondary_models import build_ours
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
gold = torch.load('tests/golden/models_golden.pt')
DEV = 'cuda'
for tag in MF.CONFIGS:
    cfg, g = MF.CONFIGS[tag], gold[tag]
    m = build_ours(cfg).to(DEV)
    if hasattr(m, 'precision'): m.precision = 'fp32'
    x = MF.inputs(cfg).to(DEV)
    m.eval()
    with torch.no_grad(): oe = m(x).cpu()
    m.train()
    out = m(x)
    (out * MF.cotangent(out.shape, cfg['seed']).to(DEV)).sum().backward()
    o = out.detach().cpu()
    de = float((oe - g['out_eval']).abs().max()); do = float((o - g['out']).abs().max())
    sc = float(g['out'].abs().max())
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    med = float(torch.tensor(list(g['grad_norms'].values())).median())
    worst = max(((abs(float(grads[n].double().norm()) - r) / max(r, 1e-3 * med)), n) for n, r in g['grad_norms'].items())
    cos = min((float(torch.nn.functional.cosine_similarity(grads[n].flatten().double(), ref.flatten().double(), dim=0)), n)
              for n, ref in g['grads'].items() if float(ref.norm()) > 1e-2 * med)
    sd = m.state_dict()
    rs = max(float(((sd[n].cpu() - ref).abs() / (ref.abs() + 1e-5)).max()) for n, ref in g['running'].items() if ref.dtype.is_floating_point)
    print("%-22s out %.2e eval %.2e (scale %.2f)  worst grad-norm dev %.2e (%s)  min cos %.6f (%s)  running %.2e" % (tag, do, de, sc, worst[0], worst[1][-40:], cos[0], cos[1][-30:], rs))
