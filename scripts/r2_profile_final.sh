# final round-2 profiles of the config-5 step and the scale-bar kernels (one GPU).  Run with gpurun; outputs in gpurun_out/.
set -x
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2f_launches_2048tiles.csv python bench.py --steps 2 --warmup 1 --device-pass-only --shards 2 > /dev/null 2> gpurun_out/r2f_l5.err
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2f_launches_256tiles.csv python bench.py --tiles 256 --steps 2 --warmup 1 --device-pass-only --shards 2 > /dev/null 2> gpurun_out/r2f_l256.err
$NCU --set full --import-source on -k regex:"k_paste_v2|k_contour_trace_slab|k_contour_hull|k_contour_measure|k_group_fused|k_containment_fused|k_measure_order" -c 9 -o gpurun_out/r2f_config5_256tiles python bench.py --tiles 256 --device-pass-only --steps 1 --warmup 1 --shards 1 --no-graph > /dev/null 2> gpurun_out/r2f_ncu5.err
$NCU --set full --import-source on -k regex:"k_hough_lines_p|k_canny|k_line_mean|k_scalebar_gray" -c 5 -o gpurun_out/r2f_scalebar python bench.py --flows-only scalebar --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/r2f_ncusb.err
ls -la gpurun_out/r2f*
