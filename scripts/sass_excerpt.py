"""profiles/r02_sass_tma_excerpt.txt: the TMA / mbarrier SASS of the fused kernels in qurious_b200/libqgpu.so
(cuobjdump -sass).  One line per distinct mnemonic and kernel.  python scripts/sass_excerpt.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "qurious_b200", "libqgpu.so")], capture_output=True, text=True).stdout
out = ["# cuobjdump -sass qurious_b200/libqgpu.so (sm_100a), excerpt: the 1-D TMA bulk copies (UBLKCP.S.G = cp.async.bulk.shared::cluster.global",
       "# .mbarrier::complete_tx::bytes) and the mbarrier operations (SYNCS.*) of the fused scan kernels and of the radix group-by kernels (k_radix_scatter_tma, k_radix_hist1_tma): one line per distinct mnemonic.",
       "# No UTMALDG (tensor-map TMA) and no UTC*MMA anywhere in the library: the path moves 1-D column tiles and contracts nothing (DESIGN 3).",
       ""]
n_tensor = len(re.findall(r"UTMALDG|UTCMMA|UTCHMMA|UTCQMMA", txt))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    seen = {}
    for l in f.split("\n"):
        m = re.search(r"(UBLKCP[.\w]*|SYNCS[.\w]*)", l)
        if m:
            e = seen.setdefault(m.group(1), [0, re.sub(r"\s+", " ", l.strip())[:140]])
            e[0] += 1
    if "UBLKCP.S.G" not in seen:
        continue
    out.append("## " + subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:160])
    for k, (c, l) in sorted(seen.items()):
        out.append(f"   x{c:<3d} {l}")
    out.append("")
out.append(f"# tensor-map TMA / tcgen05 MMA instructions in the whole library: {n_tensor}")
open(os.path.join(ROOT, "profiles", "r02_sass_tma_excerpt.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
