"""Rate of the drop-in call exactly as runners/diffpose_frame.py:365 makes it (betas and mask on the GPU, optionally a
DataParallel wrapper): time per call seen by the host loop vs wall time per call (GPU only).  NOTE: in a GPU-bound loop the
"host" figure is launch-queue back-pressure (it approaches the kernel time); the actual host cost of a call is what
tools/host_path_profile.py measures into an idle queue (17.6 us).

    python tools/api_rate.py
"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffpose_nw_b200 as D
from oracle import diffpose_oracle as O
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = D.FusedGCNdiff(D.adj_mx_from_edges(), O.default_config()).to(dev).eval()
betas = torch.from_numpy(O.beta_schedule("linear", 1e-4, 1e-3, 51)).float().to(dev)       # like the runner: on the GPU
mask = torch.ones(1, 1, 17, dtype=torch.bool, device=dev)
x = O.synthetic_poses(1024, seed=1).to(dev)
for wrap in (model, torch.nn.DataParallel(model, device_ids=[0])):
    for _ in range(20): D.generalized_steps(x, mask, range(0, 24, 12), wrap, betas, eta=0.0)
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        out = D.generalized_steps(x, mask, range(0, 24, 12), wrap, betas, eta=0.0)[0][-1]
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(type(wrap).__name__, "host %.1f us per call, wall %.1f us per call -> %.2f M poses/s through the reference-style call" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, 1024 * n / (t2 - t0) / 1e6))
