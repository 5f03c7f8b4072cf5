#!/bin/bash
# Time ONE conv layer under the kernel's experiment switches, one process per variant (the switches are read once per process).
#   scripts/probe_variants.sh SHAPE "VAR=VAL ..." "VAR=VAL ..." ...   (SHAPE = B,H,W,cin,cout,k,stride,pad; "-" = defaults)
# Prints the median duration of the tcgen05 launches of each variant (ncu gpu__time_duration, cold cache).
shape=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo "$shape-$v" | tr ' =,' '___')
  if [ "$v" = "-" ]; then envs=""; else envs="$v"; fi
  env $envs REPS=5 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_umma --csv \
      --log-file gpurun_out/probe_$tag.csv python scripts/layer_probe.py $shape - > /dev/null 2>&1
  python - "$shape" "$v" gpurun_out/probe_$tag.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 10 and r[0].isdigit()]
main = [float(r[-1]) / 1e3 for r in rows if ", 1, 0, 0>" not in r[4] and "(bool)1, (int)0" not in r[4]]
main = sorted(main[1:]) if len(main) > 1 else main
name = rows[0][4].split("(")[0] if rows else "?"
print(f"{sys.argv[1]:28s} {sys.argv[2]:40s} {main[len(main)//2] if main else float('nan'):8.1f} us  ({len(main)} launches) {name}")
PY
done
