import sys, torch
sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat
from scripts.bench_splat import smooth_flow, timeit
for c in (1, 64):
    h, w = 1152, 1920
    x = torch.randn((1, c, h, w), device="cuda"); flow = smooth_flow(h, w, 2.0, 1) + 6.5; metric = torch.randn((1, 1, h, w), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(c, round(timeit(lambda: softsplat(x, flow, metric, "soft", _variant=3), flush=flush), 4))
