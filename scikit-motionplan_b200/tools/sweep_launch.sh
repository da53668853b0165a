for wl in cfg3 cfg4_409600 cfg2_all cfg5_env_all cfg3_f64; do
  for gw in "0 0" "4 8" "4 6" "4 4" "8 8" "8 6" "8 4" "2 8" "2 6" "2 4"; do
    set -- $gw
    r=$(SKB_S2_G=$1 SKB_S2_WARPS=$2 python scikit-motionplan_b200/tools/bench_kernels.py --only $wl --reps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'configs_per_s' in d: print(round(d['frac_of_hbm'],4))
")
    echo "$wl G=$1 W=$2 -> $r"
  done
done
