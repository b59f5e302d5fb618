#!/usr/bin/env python
"""gpurun_out/r2_mma_shapes.jsonl (scripts/mma_shapes.cu on a B200) -> profiles/r2_mma_shapes.json: one compact table of
cycles per tcgen05.mma (kind::f16, K = 16, SS mode) and the conclusions drawn in DESIGN.md 4.1."""
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_mma_shapes.jsonl"
rows = [json.loads(l) for l in open(src) if l.startswith("{")]
table = {}
for r in rows:
    key = f'{r["case"]}|cg{r["cta_group"]}|M{r["M"]}|swz{r["swizzle"]}'
    if r["case"].startswith("halo"):
        key += f'|shiftA{r["a_shift_rows"]}|shiftB{r["b_shift_rows"]}'
    table.setdefault(key, {})[f'N{r["N"]}'] = r["cycles_per_mma"]
out = {
    "source": "scripts/mma_shapes.cu, one CTA (or CTA pair) per SM on all 148 SMs of a B200, clock64 around bursts of 128 / 640 "
              "back-to-back MMAs into one accumulator (difference removes the fixed latency); SM clock 1965 MHz",
    "cycles_per_mma": table,
    "model": "SS mode, cta_group::1: cycles = max((32 M + 32 N) / 128, M N / 256 * (128 / M)): operand bytes (32 B per operand row and "
             "K = 16 step, both operands) over 128 B/clk of shared-memory read bandwidth, or the math floor, whichever is larger",
    "peak_flop_per_clk_per_sm": 8192,
    "rows": rows,
}
json.dump(out, open("profiles/r2_mma_shapes.json", "w"), indent=1)
for k, v in table.items():
    print(k, v)
