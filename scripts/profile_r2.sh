#!/bin/bash
# profile_r2.sh - the round-2 ncu evidence, one GPU (run under gpurun; everything lands in gpurun_out/, digests are then
# copied to profiles/ by hand).  Numbers printed by a run under ncu are never bench values.
#   1. launch lists (gpu__time_duration.sum) of the default bench command and of the SPPM shadows line;
#   2. ncu --set full of the walk kernels of one tess-1M frame (lanes = 1, no graph: one launch per bounce level);
#   3. ncu --set full of the SPPM kernels at 1024^2 (camera / photon shade, grid build, deposit).
set -x
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# 1. launch lists
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $out/r2f_launches_default_cmd.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-optin > $out/r2f_launches_default_cmd.log 2>&1
$NCU --metrics gpu__time_duration.sum --launch-skip 600 -c 400 --csv --log-file $out/r2f_launches_sppm_shadows.csv \
    python bench.py --workload sppm-shadows-1024 --steps 4 --warmup 3 --no-cpu-baseline > $out/r2f_launches_sppm_shadows.log 2>&1
# 2. walk kernels, full set (second frame: skip the 6 traversal launches of the warm-up frame)
$NCU --set full --import-source on -k regex:"k_wh_primary|k_wh_extend|k_wh_shadow" --launch-skip 6 -c 6 -f -o /tmp/r2f_walk \
    python bench.py --steps 1 --warmup 1 --lanes 1 --graph 0 --no-sppm --no-optin --no-cpu-baseline > $out/r2f_ncu_walk.log 2>&1
bash scripts/ncu_digest.sh /tmp/r2f_walk.ncu-rep $out/r2f_ncu_walk
# 3. SPPM kernels at 1024^2, full set (one steady-state iteration: skip the first four iterations' launches)
$NCU --set full --import-source on -k regex:"k_sppm_cam_shade|k_photon_shade|k_grid_bounds|k_grid_insert|k_photon_deposit|k_sppm_update" \
    --launch-skip 64 -c 16 -f -o /tmp/r2f_sppm python scripts/sppm_bench.py --workloads sppm-shadows-1024 --variants sppm_pipeline=1 --iters 4 > $out/r2f_ncu_sppm.log 2>&1
bash scripts/ncu_digest.sh /tmp/r2f_sppm.ncu-rep $out/r2f_ncu_sppm
ls -la $out/r2f_*
