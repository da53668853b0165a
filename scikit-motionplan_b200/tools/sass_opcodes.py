"""Opcode histogram of the hot kernels, from the in-tree objects (no GPU needed):

    python scikit-motionplan_b200/tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt

For every kernel: target architecture, registers / spills from ptxas, the number of SASS instructions, the counts of the
opcodes that carry the Blackwell-specific claims of DESIGN.md (packed fp32x2 arithmetic FFMA2 / FMUL2, texture-pipe fetches TLD of the
voxel cell packs, 1-D TMA bulk stores
UBLKCP with the async-proxy fence, 256-bit global loads LDG.E.*.256 for the voxel cell packs, uniform-datapath branches,
warp shuffles), and the 25 most frequent opcodes.
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

CSRC = Path(__file__).resolve().parents[1] / "csrc"

# (object, mangled substring, what it is)
KERNELS = [
    ("skb_sphere2.o", "s2_kernelIfLi1ELi2ELi2EjLb0ELb1ELi8ELb0E", "cfg3: step kernel, fp32, CollFreeConst vs voxel grid (cell packs), all rows + Jacobian, D <= 16 (bench.py default)"),
    ("skb_sphere2.o", "s2_kernelIfLi2ELi1ELi2EjLb0ELb1ELi4ELb0E", "cfg2: step kernel, fp32, PairWiseSelfCollFreeConst, all pairs + Jacobian, D <= 8"),
    ("skb_sphere2.o", "s2_kernelIfLi1ELi1ELi1EjLb0ELb1ELi4ELb0E", "cfg1 / cfg4: step kernel, fp32, CollFreeConst vs one primitive (box), all rows + Jacobian, odd D <= 8"),
    ("skb_sphere2.o", "s2_kernelIfLi1ELi1ELi1EjLb0ELb1ELi4ELb1E", "cfg4: the same with transposed (CSC) Jacobian blocks: direct coalesced stores, no staging"),
    ("skb_sphere2.o", "s2_kernelIfLi1ELi0ELi1EyLb0ELb1ELi32ELb0E", "cfg5: step kernel, fp32, CollFreeConst vs union, all rows + Jacobian, D = 37 (64-bit masks)"),
    ("skb_sphere2.o", "s2_kernelIfLi1ELi0ELi1EyLb1ELb0ELi1ELb0E", "cfg5: step kernel, reducing mode (row minimum / validity byte), union SDF, D > 32"),
    ("skb_sphere2.o", "s2_kernelIfLi2ELi1ELi1EyLb1ELb0ELi1ELb0E", "cfg5: step kernel, pairwise lean pair scan (row minimum / validity byte), D > 32"),
    ("skb_sphere2.o", "s2_kernelIdLi1ELi2ELi2EjLb0ELb1ELi8ELb0E", "cfg3 fp64 build option: step kernel, double"),
    ("skb_pose.o", "pose_kernelIfLi2E", "cfg5: pose kernel, fp32, XYZW rows"),
    ("skb_pose.o", "pose_kernelIfLi1E", "cfg4: pose kernel, fp32, RPY rows"),
    ("skb_sphere.o", "sphere_kernelIfLi0ELb1ELb0ELi1ELi2Ej", "KinematicsMap.map: sphere kernel, fp32, point features + Jacobian"),
    ("skb_com.o", "com_kernelIf", "COMStabilityConst: centre-of-mass kernel, fp32"),
]
WATCH = ["FFMA2", "FMUL2", "FFMA", "UBLKCP", "FENCE", "TLD", "TEX", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BRA.U", "BRA", "CALL", "STL", "LDL", "DFMA"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(CSRC / obj)], capture_output=True, text=True).stdout
    cur, table, arch = None, {}, ""
    for ln in out.splitlines():
        m = re.match(r"\s*arch = (\S+)", ln)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            table[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            table[cur].append(m.group(2))
    return table, arch


def ptxas_stats(obj):
    log = CSRC / ("ptxas_" + obj.replace("skb_", "").replace(".o", "") + ".log")
    stats, cur = {}, None
    if log.exists():
        for ln in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                stats.setdefault(cur, {})["spill"] = f"stack {m.group(1)} B, spill stores {m.group(2)} B, spill loads {m.group(3)} B"
            m = re.search(r"Used (\d+) registers", ln)
            if m and cur:
                stats.setdefault(cur, {})["regs"] = m.group(1)
    return stats


def main():
    cache = {}
    print("# SASS opcode histogram of the hot kernels (cuobjdump -sass over scikit-motionplan_b200/csrc/*.o; tools/sass_opcodes.py)\n")
    for obj, sub, what in KERNELS:
        if obj not in cache:
            cache[obj] = (functions(obj), ptxas_stats(obj))
        (table, arch), stats = cache[obj]
        names = [n for n in table if sub in n]
        if not names:
            print(f"## {what}\n   (no function matching {sub} in {obj})\n")
            continue
        name = names[0]
        ops = table[name]
        c = Counter(ops)
        st = stats.get(name, {})
        print(f"## {what}")
        print(f"   {obj} :: {name}")
        print(f"   arch {arch}; {st.get('regs', '?')} registers; {st.get('spill', 'n/a')}; {len(ops)} SASS instructions")
        fam = Counter()
        for op, n in c.items():
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    fam[w] += n
                    break
        print("   families: " + ", ".join(f"{w} {fam[w]}" for w in WATCH if fam[w]))
        wide = {op: n for op, n in c.items() if op.startswith("LDG") and "256" in op}
        if wide:
            print("   256-bit global loads: " + ", ".join(f"{k} {v}" for k, v in sorted(wide.items())))
        bulk = {op: n for op, n in c.items() if op.startswith("UBLKCP") or op.startswith("FENCE")}
        if bulk:
            print("   TMA bulk stores / async-proxy fence: " + ", ".join(f"{k} {v}" for k, v in sorted(bulk.items())))
        print("   top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(25)))
        print()


if __name__ == "__main__":
    main()
