#!/bin/bash
# One GPU-box pass that produces the evidence under profiles/: GPU suite, bench line, the
# gpu__time_duration launch list of the bench command and `ncu --set full` captures of the
# scan kernel, the refinement kernel and the rotation scan.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile_round.sh r02f'
# Every ncu pass runs only after the same command exited 0 without the profiler.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
set -x
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_gputest.log 2>&1; echo "gputest rc=$?" >> $out/${tag}_gputest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 300 $B > $out/${tag}_bench_short.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^grid_kernel" -s 3 -c 1 \
    -o $out/${tag}_grid -f $B > $out/${tag}_ncu_grid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^refine_kernel" -s 3 -c 1 \
    -o $out/${tag}_refine -f $B > $out/${tag}_ncu_refine.log 2>&1
K="python scripts/gpu_ncu_kinds.py"
timeout 300 $K > $out/${tag}_kinds.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^grid_kernel" -c 4 \
    -o $out/${tag}_kinds -f $K > $out/${tag}_ncu_kinds.log 2>&1
# the reports exceed what gpurun copies back: export the pages here, keep only the CSVs
for r in grid refine kinds; do
  f=$out/${tag}_$r.ncu-rep
  [ -f $f ] || continue
  ncu -i $f --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
  ncu -i $f --page details --csv > $out/${tag}_${r}_details.csv 2>/dev/null
  ncu -i $f --page source --csv > $out/${tag}_${r}_source.csv 2>/dev/null
  rm -f $f
done
ls -la $out
