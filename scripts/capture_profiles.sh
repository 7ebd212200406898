#!/bin/bash
# ncu captures behind profiles/: run under gpurun (one GPU), then `python scripts/summarize_profiles.py rNN` here.
#   launch lists: cold-cache, serialised per-launch durations of 4 frames per config
#   full captures: every frame kernel of the first frames (the summary uses each kernel's last launch), source-level counters included
mkdir -p gpurun_out
for w in ${@:-C1 C2 C3 C4}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv \
      python scripts/profile_run.py --workload $w --frames 4 > gpurun_out/prof_$w.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k 'regex:geom_kernel|cull_kernel|vertex_kernel|geom_list_kernel|clip_kernel|mid_kernel|sort_big_kernel|bin_|lean_resolve_kernel|tile_kernel|shade_kernel' \
      -c 20 -f -o gpurun_out/full_$w python scripts/profile_run.py --workload $w --frames 3 --opt graphs=0 >> gpurun_out/prof_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
