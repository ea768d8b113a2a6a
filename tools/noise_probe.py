"""Debug: gradient differences eager vs eager and CUDA-graph replay vs eager for one ConsecutiveSwinBlocks pair."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pwa_b200
from pwa_b200.graphs import GraphedStep
DEV = "cuda:0"
torch.manual_seed(7)
pair = pwa_b200.ConsecutiveSwinBlocks(hidden_channels=48, num_heads=4, pos_bias_embed_dim=64, max_prompts=1,
                                      tokens_per_prompt=64, window_size=(8, 8, 4), down=True).to(DEV)
prompts = [torch.nn.Parameter(0.2 * torch.randn(64, 48, device=DEV)) for _ in range(2)]
params = list(pair.parameters()) + prompts
names = [n for n, _ in pair.named_parameters()] + ["p0", "p1"]
def step(x):
    p = tuple(t.to(x.dtype).unsqueeze(0).expand(x.shape[0], -1, -1) for t in prompts)
    loss = pair(x, p).float().square().mean(); loss.backward(); return loss.detach()
xs = [torch.randn(2, 48, 16, 16, 8, device=DEV).bfloat16() for _ in range(2)]
def eager(x):
    for p in params: p.grad = None
    xe = x.clone().requires_grad_(True)
    l = step(xe)
    return l.clone(), xe.grad.clone(), [p.grad.clone() for p in params]
e0 = eager(xs[1]); e1 = eager(xs[1])
g = GraphedStep(step, [xs[0].clone().requires_grad_(True)], params)
lg = g(xs[1]).clone(); gg = [p.grad.clone() for p in params]; xg = g.input_grads[0].clone()
lg2 = g(xs[1]).clone(); gg2 = [p.grad.clone() for p in params]
print("loss eager", e0[0].item(), "graph", lg.item(), "diff", (e0[0] - lg).abs().item())
print("xgrad diff", (e0[1].float() - xg.float()).abs().max().item(), "max", e0[1].float().abs().max().item())
for n, a, b, c, d in zip(names, e0[2], e1[2], gg, gg2):
    print(f"{n:42s} max {a.abs().max().item():.3e} ee {(a-b).abs().max().item():.2e} ge {(a-c).abs().max().item():.2e} gg {(c-d).abs().max().item():.2e}")
