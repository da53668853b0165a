#!/bin/bash
# Turns the gpurun_out/prof_* files of tools/refresh_profiles.sh into the tracked profiles/<tag>_* set (run here, where the
# pulled CSV exports are).   bash scikit-motionplan_b200/tools/postprocess_profiles.sh r1_k
set -eu
TAG=${1:?tag, e.g. r1_k}
T=scikit-motionplan_b200/tools; P=profiles; O=gpurun_out
LIB=scikit-motionplan_b200/skmp_b200/libskmp_b200.so
W=$(mktemp -d)
(cd $W && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null && nvdisasm -g -c skb_sphere2.sm_100a.cubin > s2.sass)
cp $O/prof_bench.json $P/${TAG}_bench.json
cp $O/prof_bench_reference_arm.json $P/${TAG}_bench_reference_arm.json
[ -s $O/prof_bench_10m.json ] && cp $O/prof_bench_10m.json $P/${TAG}_bench_10m.json
for w in cfg2 cfg4 cfg5; do [ -s $O/prof_bench_$w.json ] && cp $O/prof_bench_$w.json $P/${TAG}_bench_$w.json; done
grep -v '^\[skb\]' $O/prof_kernels.jsonl > $P/${TAG}_kernels.jsonl
cp $O/prof_launches.csv $P/${TAG}_launches.csv
[ -s $O/prof_ik.json ] && cp $O/prof_ik.json $P/${TAG}_ik.json
for m in cfg3:collfree_cfg3 cfg2_all:pairwise_cfg2 cfg4_409600:collfree_cfg4 cfg5_env_all:collfree_jaxon_env \
         cfg5_self_valid:pairwise_jaxon_validity cfg5_pose:pose cfg5_com:com; do
    python $T/ncu_summary.py $O/prof_${m%%:*}_raw.csv > $P/${TAG}_${m##*:}_ncu_full.json
done
python $T/ncu_sections.py $O/prof_cfg3_src.csv $W/s2.sass fLi1ELi2ELi2EjLb0ELb1ELi8ELb0EE 1250000 > $P/${TAG}_collfree_cfg3_sections.txt
python $T/ncu_sections.py $O/prof_cfg4_409600_src.csv $W/s2.sass fLi1ELi1ELi1EjLb0ELb1ELi4ELb0EE 409600 > $P/${TAG}_collfree_cfg4_sections.txt
python $T/ncu_sections.py $O/prof_cfg5_env_all_src.csv $W/s2.sass fLi1ELi0ELi1EyLb0ELb1ELi32ELb0EE 1000000 > $P/${TAG}_collfree_jaxon_env_sections.txt
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.load(open(f"profiles/{tag}_collfree_cfg3_ncu_full.json"))[0]
def b(k):
    v = d[k]
    return v["value"] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v["unit"]]
r, w = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
t = json.load(open("profiles/traffic_collfree_cfg3.json"))
t.update(kernel=d["kernel"], dram_bytes_read=r, dram_bytes_write=w, dram_bytes_per_launch=r + w,
         ratio=(r + w) / t["algorithmic_bytes_per_launch"],
         source=f"profiles/{tag}_collfree_cfg3_ncu_full.json (ncu --set full --clock-control none, one launch of bench.py's timed step)")
json.dump(t, open("profiles/traffic_collfree_cfg3.json", "w"), indent=1)
PY
rm -rf $W
