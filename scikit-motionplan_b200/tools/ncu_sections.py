"""Group an ncu source-page CSV of skb::s2_kernel by kernel section (line ranges of skb_sphere2.cu / skb_device.cuh)
and print each section's share of executed warp instructions and stall samples.

  cuobjdump -xelf all libskmp_b200.so ; nvdisasm -g -c skb_sphere2.sm_100a.cubin > s2.sass
  ncu -i prof.ncu-rep --page source --csv > prof_src.csv
  python ncu_sections.py prof_src.csv s2.sass <mangled-kernel-substring> <configs-per-launch>
"""
import csv
import re
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ncu_lines  # noqa: E402

SRC = Path(__file__).resolve().parents[1] / "csrc"


def anchors():
    """section boundaries from marker strings, so the table survives edits of the kernel source"""
    s2 = (SRC / "skb_sphere2.cu").read_text().splitlines()
    dv = (SRC / "skb_device.cuh").read_text().splitlines()

    def find(lines, pat, start=0):
        for i in range(start, len(lines)):
            if pat in lines[i]:
                return i + 1
        raise KeyError(pat)

    a = {}
    a["fma2_lo"], a["fma2_hi"] = find(s2, "__device__ __forceinline__ float2 fma2"), find(s2, "enum { SDF_UNION")
    a["pair_lo"], a["pair_hi"] = find(s2, "template <typename T> struct Pair2;"), find(s2, "the corner-pack fast path exists for fp32 only")
    a["grid_lo"], a["grid_hi"] = find(s2, "struct GridPend {"), find(s2, "the corner-pack fast path exists for fp32 only")
    a["fsplit_lo"], a["fsplit_hi"] = find(s2, "auto front_issue = "), find(s2, "// twist pair x wrench for the columns")
    a["prologue_lo"], a["prologue_hi"] = find(s2, "extern __shared__"), find(s2, "auto fetch_q")
    a["q_lo"], a["q_hi"] = find(s2, "auto fetch_q"), find(s2, "phase 1: frames + twists")
    a["front_lo"], a["front_hi"] = find(s2, "auto front = "), find(s2, "if (!RED) {")
    a["step_lo"] = find(s2, "if (!RED) {")
    a["contract_lo"], a["contract_hi"] = find(s2, "auto contract = "), find(s2, "auto body = ")
    a["body_lo"] = find(s2, "auto body = ")
    a["jac_lo"] = find(s2, "if (lane == 0) bulk_wait_read<0>();")
    a["ship_lo"] = find(s2, "// ship: the dense")
    a["ship_hi"] = find(s2, "if (cntB > 0) for (int c = 0;")
    a["fk_lo"], a["fk_hi"] = find(dv, "void fk_phase("), find(dv, "bulk async store helpers")
    a["real_hi"] = find(dv, "Fixed-association building blocks")
    a["wrench_hi"] = find(dv, "SDF primitives (local frame)")
    return a


def main():
    src_csv, sass, kernel, n_cfg = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    A = anchors()

    def grp(f, l):
        if f is None:
            return "unattributed"
        if f == "skb_device.cuh":
            if A["fk_lo"] <= l < A["fk_hi"]:
                return "phase 1: FK walk (3 lanes per node)"
            if l < A["wrench_hi"]:
                return "wrench moment (explicit mul/fma)"
            if l < A["fk_lo"]:
                return "SDF primitives (box / cylinder / sphere / generic grid, skb_device.cuh)"
            return "bulk-store / fence helpers"
        if f == "skb_sphere2.cu":
            if A["fma2_lo"] <= l < A["fma2_hi"] or A["pair_lo"] <= l < A["pair_hi"]:
                return "phase 2: packed FFMA2 chains"
            if A["grid_lo"] <= l < A["grid_hi"]:
                return "grid SDF (gather + trilinear + gradient)"
            if A["prologue_lo"] <= l < A["prologue_hi"]:
                return "kernel-parameter / shared-base re-materialisation (attributed to the prologue lines)"
            if A["q_lo"] <= l < A["q_hi"]:
                return "q fetch + sin/cos pass"
            if A["front_lo"] <= l < A["front_hi"]:
                return "phase 2: sphere centre, value, masks"
            if A["fsplit_lo"] <= l < A["fsplit_hi"]:
                return "phase 2: sphere centre, value, masks"
            if A["contract_lo"] <= l < A["contract_hi"]:
                return "phase 2: column-mask selects"
            if A["step_lo"] <= l < A["contract_lo"]:
                return "phase 2: step header + lane tables"
            if A["body_lo"] <= l < A["jac_lo"]:
                return "phase 2: value stores"
            if A["jac_lo"] <= l < A["ship_lo"]:
                return "phase 2: twist loads, row stores, loop control"
            if A["ship_lo"] <= l < A["ship_hi"]:
                return "phase 2: fence + TMA bulk store"
            return "other (skb_sphere2.cu)"
        if "intrinsics" in f:
            return "warp shuffles / syncwarp"
        return f

    amap = ncu_lines.sass_lines(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ia, ie, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    base, tot, stot = None, 0, 0
    groups = defaultdict(lambda: [0, 0])
    for r in rows[2:]:
        if len(r) <= ist or not r[ie].isdigit():
            continue
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = addr
        key = amap.get(addr - base, ((None, 0), ""))[0]
        n, s = int(r[ie]), int(r[ist] or 0)
        g = grp(*key) if key else "unattributed"
        groups[g][0] += n
        groups[g][1] += s
        tot += n
        stot += s
    print(f"kernel {kernel}: {tot} warp instructions = {tot / n_cfg:.0f} per configuration, {stot} stall samples")
    for g, (n, s) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * n / tot:6.2f}% inst ({n / n_cfg:7.1f} / cfg) {100 * s / max(1, stot):6.2f}% stall   {g}")


if __name__ == "__main__":
    main()
