import torch, sys
sys.path.insert(0, ".")
import diffpose_nw_b200 as D
from oracle import diffpose_oracle as O
dev = torch.device("cuda:0"); torch.set_grad_enabled(False)
torch.manual_seed(0)
m = D.FusedGCNdiff(D.adj_mx_from_edges(), O.default_config()).to(dev).eval()
x = O.synthetic_poses(4096, seed=1).to(dev)
t = torch.randint(0, 50, (4096,), device=dev).float()
for eng in ("auto", "tcg", "fp32"):
    m.set_engine(eng)
    for _ in range(3): m(x, None, t, 0)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30 if eng != "fp32" else 3
    e0.record()
    for _ in range(n): out = m(x, None, t, 0)
    e1.record(); torch.cuda.synchronize()
    print(f"GCNdiff.forward, 4096 samples with per-sample t, engine {eng} -> {m.forward_engine()}: {e0.elapsed_time(e1) / n:.4f} ms per call")
