"""Device time of each of the first steps after a synchronisation (what a 20-step timed window sees):
events around every call of configs[1]."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffpose_nw_b200 as D
from oracle import diffpose_oracle as O
dev = torch.device("cuda:0")
torch.manual_seed(0); torch.set_grad_enabled(False)
model = D.FusedGCNdiff(D.adj_mx_from_edges(), O.default_config()).to(dev).eval()
betas = torch.from_numpy(O.beta_schedule("linear", 1e-4, 1e-3, 51)).float()
x = O.synthetic_poses(1024, seed=1).to(dev)
seq = range(0, 24, 12)
steps = D.ddim_steps(betas, seq, 0.0)
for _ in range(200):
    D.sample(model, x, None, seq, betas, steps=steps)
for trial in range(3):
    torch.cuda.synchronize()
    time.sleep(0.05 * trial)
    n = 24
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(n):
        D.sample(model, x, None, seq, betas, steps=steps)
        ev[i + 1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n)]
    print(f"trial {trial} (idle {50 * trial} ms before): host enqueue {1e6 * (t1 - t0) / n:.1f} us/call; per-step us:", " ".join(f"{v:.0f}" for v in d), f"| total {ev[0].elapsed_time(ev[n]) * 1e3:.0f} us")
# the same with one untimed lead-in step between the sync and the first event
for trial in range(2):
    torch.cuda.synchronize()
    D.sample(model, x, None, seq, betas, steps=steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        D.sample(model, x, None, seq, betas, steps=steps)
    e1.record()
    torch.cuda.synchronize()
    print(f"lead-in step, 20 timed steps: {e0.elapsed_time(e1) * 1e3 / 20:.2f} us per step")
for trial in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        D.sample(model, x, None, seq, betas, steps=steps)
    e1.record()
    torch.cuda.synchronize()
    print(f"no lead-in, 20 timed steps: {e0.elapsed_time(e1) * 1e3 / 20:.2f} us per step")
