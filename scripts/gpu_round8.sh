#!/bin/bash
# ncu --set full of the two gather kernels on the gentle flow (first variant-4 call, first variant-3 call)
cat > /tmp/one_splat.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat
c, h, w = 64, 1152, 1920
g = torch.Generator(device="cpu").manual_seed(1)
lo = 2.0 * torch.randn((1, 2, h // 16, w // 16), generator=g)
flow = (torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False) + 6.5).cuda()
x = torch.randn((1, c, h, w), device="cuda"); metric = torch.randn((1, 1, h, w), device="cuda")
for v in (4, 3):
    softsplat(x, flow, metric, "soft", _variant=v); torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:splat_gather -o gpurun_out/r2_splat_gather_full -f python /tmp/one_splat.py > /dev/null 2>&1
ncu -i gpurun_out/r2_splat_gather_full.ncu-rep --page raw --csv > gpurun_out/r2_splat_gather_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_splat_gather_full*
