"""GAE pipelined-scan ring-shape sweep (T=512 x 65,536 envs): CUDA-graph replay of 10 launches per variant."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import marl_sat_b200 as M            # noqa: E402
from marl_sat_b200 import _lib       # noqa: E402
lib = _lib.load()
dev = torch.device("cuda", 0)
T, B = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device=dev).manual_seed(0)
reward = (torch.rand((T, B), generator=g, device=dev) < 0.01).float()
done = (torch.rand((T, B), generator=g, device=dev) < 0.005).to(torch.uint8)
value = torch.randn((T, B), generator=g, device=dev)
last = torch.randn((B,), generator=g, device=dev)
adv, tgt = torch.empty((T, B), device=dev), torch.empty((T, B), device=dev)
stats = torch.zeros(3, dtype=torch.float64, device=dev)
names = {0: "pipelined, width by batch", 4: "pipelined, 4 columns/lane", 2: "pipelined, 2 columns/lane",
         1: "pipelined, 1 column/lane", -1: "register-chunked / segmented"}
if len(sys.argv) > 2:
    lib.msat_tune(b"gae_pipe_min_cols", int(sys.argv[2]))
for v in [-1, 0, 4, 2, 1]:
    lib.msat_tune(b"gae_plain", 1 if v < 0 else 0)
    lib.msat_tune(b"gae_variant", max(v, 0))
    for _ in range(3):
        M.calculate_gae(reward, done, value, last, 0.995, 0.95, stats=stats, out=(adv, tgt))
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            M.calculate_gae(reward, done, value, last, 0.995, 0.95, stats=stats, out=(adv, tgt))
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"variant {v:2d} {names[v]:26s} {best*1e3:8.1f} us  {17.0*T*B/(best*1e-3)/1e9:7.0f} GB/s  frac {17.0*T*B/(best*1e-3)/1e9/6546.6:.3f}")
lib.msat_tune(b"gae_plain", 0); lib.msat_tune(b"gae_variant", 0)
