"""profiles/r2_sass.md from `cuobjdump -sass simpleicp_b200/libsicp_b200.so` (build host, no GPU)."""
import re
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
txt = subprocess.run(["cuobjdump", "-sass", str(REPO / "simpleicp_b200" / "libsicp_b200.so")],
                     capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
out = ["# SASS evidence (round 2): `cuobjdump -sass simpleicp_b200/libsicp_b200.so`, sm_100a cubin",
       "",
       "Generated on the build host by `tools/sass_excerpt.py` (no GPU needed). Counts of the mnemonics that",
       "prove the Blackwell-specific paths, per kernel, then the TMA / mbarrier lines of `k_bf_nn` verbatim.",
       "",
       "| kernel | instructions | UBLKCP (1-D TMA bulk copy) | SYNCS (mbarrier) | LDG.E.*.256 (32-byte record loads) | DFMA/DADD/DMUL | ATOM/RED | SHFL | LDL/STL (local memory) |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
keep = ("k_bf_nn", "k_match_grid_coop", "k_match_batch", "k_rs_fused", "k_rs_batch", "k_reject_solve", "k_knn",
        "k_pca", "k_transform", "k_match_grid")
bf = None
rsf = None
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("(anonymous namespace)::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", dem).split("::")[-1]
    if not short.startswith(keep):
        continue
    lines = [ln for ln in f.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/", ln)]
    ins = [re.sub(r".*/\*[0-9a-f]{4,5}\*/\s+", "", ln).split(";")[0].strip() for ln in lines]
    cnt = lambda p: sum(1 for i in ins if re.search(p, i))  # noqa: E731
    out.append(f"| `{short}` | {len(ins)} | {cnt(r'UBLKCP')} | {cnt(r'SYNCS')} | {cnt(r'LDG\.E\.[A-Z0-9.]*256')} | "
               f"{cnt(r'^@?!?P?[0-9T]* ?D(FMA|ADD|MUL)|\bD(FMA|ADD|MUL)\b')} | {cnt(r'\b(ATOM|ATOMS|ATOMG|RED|REDG|REDS)\b')} | "
               f"{cnt(r'SHFL')} | {cnt(r'\b(LDL|STL)\b')} |")
    if short.startswith("k_bf_nn"):
        bf = [ln for ln in lines if re.search(r"UBLKCP|SYNCS|FENCE|MBAR", ln)]
    if short == "k_rs_fused":
        rsf = [ln for ln in lines if re.search(r"UBLKCP|SYNCS|FENCE", ln)]
out += ["", "## `k_bf_nn`: the TMA / mbarrier instructions", "", "```"]
out += [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in (bf or [])][:40] + ["```", ""]
out += ["## `k_rs_fused`: the three bulk copies that stage a block's share of the moment pass", "", "```"]
out += [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in (rsf or [])][:20] + ["```", ""]
(REPO / "profiles" / "r2_sass.md").write_text("\n".join(out))
print("\n".join(out))
