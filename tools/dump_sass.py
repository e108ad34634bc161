"""Write compact SASS listings (opcodes + operands, no encodings) of the hot
kernels into profiles/sass/ -- evidence of what the compiler emitted
(cp.async = LDGSTS, LDS.128, FFMA, DADD/SHFL for the wavefront, UTC*MMA once the
tcgen05 path lands)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "abnet3_b200", "libabnet3_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KERNELS = {
    "align_class_kernel_4_4": "_ZN3abn18align_class_kernelILi4ELi4EEEvNS_9AlignArgsE",
    "align_long_kernel": "_ZN3abn17align_long_kernelENS_9AlignArgsEi",
    "pair_loss_kernel": "_ZN3abn16pair_loss_kernelEPKfS1_S1_liiffPfS2_S2_",
}
os.makedirs(OUT, exist_ok=True)
for name, sym in KERNELS.items():
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
    lines = []
    for ln in txt.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append("%s  %s" % (m.group(1), re.sub(r"\s+", " ", m.group(2)).strip()))
    ops = {}
    for ln in lines:
        op = ln.split("  ", 1)[1].split(" ")
        op = op[1] if op[0].startswith("@") else op[0]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    head = "# %s (%s): %d instructions; opcode histogram: %s\n" % (
        name, sym, len(lines), ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
    with open(os.path.join(OUT, name + ".sass"), "w") as fh:
        fh.write(head + "\n".join(lines) + "\n")
    print(head.strip())
