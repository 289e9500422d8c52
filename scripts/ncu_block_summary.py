"""Summarise the ncu launch list of ONE replay of the captured bilevel block (scripts/ncu_graph_block.py) by kernel family:
   python scripts/ncu_block_summary.py <launches.csv> <out.md>"""
import collections, csv, sys
path, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
idx = {h: j for j, h in enumerate(hdr)}
FAMILIES = [("lds::k2_mma_kernel", "K2 `k2_mma_kernel` (tcgen05 propagate, every product with a sampled graph)"),
            ("lds::k2_prep_kernel", "K2 operand pack `k2_prep_kernel` (fp32 -> bf16 hi/lo, K-major)"),
            ("lds::k1_sample_kernel", "K1 `k1_sample_kernel<false,true>` (device-side Philox step)"),
            ("lds::k3_tc_kernel", "K3+K4 `k3_tc_kernel` (all factor pairs of the block, one pass over theta)"),
            ("lds::k3_pack_kernel", "K3 factor pack `k3_pack_kernel`"),
            ("lds::spmm_csr_kernel", "CSR feature products `spmm_csr_kernel`"),
            ("lds::row_linear_kernel", "skinny X W^T `row_linear_kernel`"),
            ("lds::gram_tn", "skinny A^T B / column sums `gram_tn_cluster_kernel` / `gram_tn_kernel`"),
            ("lds::adam_step", "differentiable Adam step and its backward `adam_step(_backward)_kernel`"),
            ("lds::masked_nll", "masked NLL on the logits: forward, gradient, gradient of the gradient `masked_nll_*_kernel`"),
            ("lds::row_dot2", "r-gradient of the normalised propagation `row_dot2_kernel`"),
            ("gemm", "cuBLAS / CUTLASS GEMMs (skinny dense products left in torch)"),
            ("reduce_kernel", "torch reductions"), ("elementwise", "torch elementwise"), ("", "other torch kernels (softmax, nll, index, cat, copy, dropout)")]
agg = collections.OrderedDict((label, []) for _, label in FAMILIES)
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]]) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(r[idx["Metric Unit"]], 1)
    name = r[idx["Kernel Name"]]
    for key, label in FAMILIES:
        if key in name:
            agg[label].append(v)
            break
total = sum(sum(v) for v in agg.values())
count = sum(len(v) for v in agg.values())
lines = ["# Round 1 (snapshot m) — ncu launch list of ONE replay of the captured bilevel block, Citeseer shape", "",
         "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/ncu_graph_block.py citeseer`",
         "(raw list: `r01m_launches_graph_block.csv`). One block = 5 `inner_opt_step` + 1 `hyper_opt_step` with the hypergradient through the 5",
         "unrolled steps (src/trainers/bilevel.py:53-73). ncu serialises and cold-starts every launch, so the absolute sum is above the",
         "3.05 ms the replay takes when timed alone; the SHARES are what to read.", "",
         "| kernel family | launches | mean us | total us | share |", "|---|---|---|---|---|"]
for label, v in agg.items():
    if v:
        lines.append(f"| {label} | {len(v)} | {sum(v) / len(v):.2f} | {sum(v):.1f} | {100 * sum(v) / total:.1f}% |")
lines += ["", f"{count} launches, {total / 1e3:.3f} ms summed kernel time."]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
