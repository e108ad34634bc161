"""Write compact SASS listings (opcodes + operands, no encodings) of the hot kernels into
profiles/sass/ -- evidence of what the compiler emitted: UTCHMMA / LDTM / UTMALDG / UTMASTG in the
tcgen05 kernels, UBLKCP (1-D TMA) + FFMA + LDS.128 in the distance kernels, DADD / SHFL in the DTW
wavefronts, red / st to peer memory in the exchange kernel.

    python tools/dump_sass.py            (after python -m abnet3_b200.build)
"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "abnet3_b200", "libabnet3_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# output name -> regex on the demangled-ish mangled symbol
WANT = {
    "align_class_kernel_4_4": r"align_class_kernelILi4ELi4E",
    "align_stack_kernel_4_4": r"align_stack_kernelILi4ELi4E",
    "dtw_skew_kernel_2": r"dtw_skew_kernelILi2E",
    "long_tile_kernel_stacked_4": r"long_tile_kernelILb1ELi4E",
    "long_tile_kernel_generic_4": r"long_tile_kernelILb0ELi4E",
    "dtw_band_kernel": r"dtw_band_kernel",
    "mlp_chain_kernel_0_forward": r"mlp_chain_kernelILi0ELb0ELb0E",
    "mlp_chain_kernel_0_forward_loss": r"mlp_chain_kernelILi0ELb0ELb1E",
    "mlp_chain_kernel_1_dgrad": r"mlp_chain_kernelILi1ELb0ELb0E",
    "mlp_chain_kernel_0_forward_dropout": r"mlp_chain_kernelILi0ELb1ELb0E",
    "tc_group_kernel_256_2": r"tc_group_kernelILi256ELi2E",
    "gather_bf16_kernel": r"gather_bf16_kernel",
    "pair_loss_dz_vec_kernel": r"pair_loss_dz_vec_kernelILb0E",
    "optimizer_fused_kernel": r"optimizer_fused_kernel",
    "dp_push_kernel": r"14dp_push_kernel",
    "dp_push1_kernel": r"15dp_push1_kernel",
}
os.makedirs(OUT, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, name = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        funcs[name] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and name:
        funcs[name].append("%s  %s" % (m.group(1), re.sub(r"\s+", " ", m.group(2)).strip()))
for out, pat in WANT.items():
    hits = [f for f in funcs if re.search(pat, f)]
    if not hits:
        print("missing:", out)
        continue
    sym, lines = hits[0], funcs[hits[0]]
    ops = collections.Counter()
    for ln in lines:
        op = ln.split("  ", 1)[1].split(" ")
        op = op[1] if op[0].startswith("@") else op[0]
        ops[op.split(".")[0]] += 1
    head = "# %s (%s): %d instructions; opcode histogram: %s\n" % (
        out, sym, len(lines), ", ".join("%s %d" % kv for kv in ops.most_common(16)))
    with open(os.path.join(OUT, out + ".sass"), "w") as fh:
        fh.write(head + "\n".join(lines) + "\n")
    print(head.strip()[:240])
