"""profiles/r02_sass_resource_usage.txt: what the sm_100a build of libila_b200.so actually contains — per kernel the ptxas
resource usage (cuobjdump --dump-resource-usage) and a SASS mnemonic histogram (cuobjdump -sass), plus the ELF arch line.
Runs without a GPU."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "infinitelinearalgebra.jl_b200", "libila_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_resource_usage.txt")
HOT = ["ql_toeplitz_l11_kernelILi4ELb0ELb0", "qr_forward_trio_kernelILi1E", "qr_forward_kernelILi1ELi1ELb0ELb1", "qr_backsolve_kernelILi1ELi1ELb0ELb1ELb1ELi16E", "chol_forward_kernelILi2",
       "ql_blocktri_l11_kernelIdLi2", "triconj_bands_kernel", "qr_blockbanded_kernel", "qr_almostbanded_kernel", "ql_pert_head_kernelILb0"]
ru = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()[:150]
lines = []
arch = sorted(set(re.findall(r"arch = (sm_\w+)", ru)))
lines.append(f"# libila_b200.so — cuobjdump evidence (commit {subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()})")
lines.append(f"# embedded cubin architectures: {', '.join(arch)}   (built with -gencode arch=compute_100a,code=sm_100a)")
res = {}
cur = None
for l in ru.splitlines():
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in l:
        res[cur] = l.strip()
        cur = None
hist = {}
cur = None
for l in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m and cur:
        hist[cur][m.group(1)] += 1
lines.append(f"# {len(res)} kernels / device functions in the library\n")
for key in HOT:
    for fn in sorted(res):
        if key in fn:
            h = hist.get(fn, collections.Counter())
            tot = sum(h.values())
            fam = collections.Counter()
            for k, v in h.items():
                fam[k.split(".")[0]] += v
            lines.append(f"## {demangle(fn)}")
            lines.append(f"   {res[fn]}")
            lines.append(f"   SASS instructions: {tot};  top mnemonics: " + ", ".join(f"{k} {v}" for k, v in fam.most_common(14)))
            special = {k: v for k, v in h.items() if any(t in k for t in ("MUFU", "UBLKCP", "UTMA", "SYNCS", "LDGSTS", "BAR", "SHFL", "DMMA", "LDS", "STS", "CCTL", "ELECT"))}
            lines.append("   of note: " + ", ".join(f"{k} {v}" for k, v in sorted(special.items())))
            lines.append("")
            break
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
