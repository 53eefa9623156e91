# round-2 profiles (one GPU): launch lists + ncu --set full captures.  Run with gpurun; outputs in gpurun_out/.
set -x
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_2048tiles_device_pass.csv python bench.py --steps 2 --warmup 1 --device-pass-only --shards 2 > gpurun_out/r2_l5.out 2> gpurun_out/r2_l5.err
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_config3a.csv python bench.py --flows-only config3a --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r2_l3a.err
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_config4.csv python bench.py --flows-only config4 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r2_l4.err
$NCU --set full --import-source on -k regex:"k_morph|k_overlap_first_come|k_resize_nearest_place|k_sp_rank|k_sp_pairs|k_column_gate|k_sp_resolve" -c 14 -o gpurun_out/r2_flows_config3b python bench.py --flows-only config3b --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r2_ncu3b.err
$NCU --set full --import-source on -k regex:"k_paste_v2|k_contour_trace_slab|k_contour_hull|k_contour_measure|k_group_fused|k_containment_fused" -c 7 -o gpurun_out/r2_config5_256tiles python bench.py --tiles 256 --device-pass-only --steps 1 --warmup 1 --shards 1 --no-graph > /dev/null 2> gpurun_out/r2_ncu5.err
ls -la gpurun_out/*.ncu-rep
