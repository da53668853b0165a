"""Join an ncu source-page CSV (per-SASS-address counters) with nvdisasm -g line info, and print the
per-CUDA-source-line share of executed instructions and stall samples for one kernel.

  cuobjdump -xelf all libskmp_b200.so ; nvdisasm -g -c skb_kernels.sm_100a.cubin > k.sass
  ncu -i prof.ncu-rep --page source --csv > prof_src.csv
  python ncu_lines.py prof_src.csv k.sass <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """address(hex int) -> (file, line) for the given kernel's .text section"""
    out, cur, active = {}, None, False
    for ln in open(path, errors="ignore"):
        if ln.startswith(".section") or ".text." in ln and ln.strip().startswith(".section"):
            active = kernel in ln
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and active:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = sass_lines(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ia, ie, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    base = None
    per = defaultdict(lambda: [0, 0])
    tot = stot = 0
    for r in rows[2:]:
        if len(r) <= ist or not r[ie].isdigit():
            continue
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = addr
        key = amap.get(addr - base, ((None, 0), ""))[0]
        n, s = int(r[ie]), int(r[ist] or 0)
        per[key][0] += n
        per[key][1] += s
        tot += n
        stot += s
    print(f"kernel {kernel}: {tot} warp instructions, {stot} stall samples")
    for key, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * n / tot:6.2f}% inst  {100.0 * s / max(1, stot):6.2f}% stall  {key}")


if __name__ == "__main__":
    main()
