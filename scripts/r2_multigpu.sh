set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -c 300 gpurun_out/r2_bench_8gpu.err
$TR --nproc-per-node 8 --master-port 29702 bench.py --workload config3a --steps 5 --warmup 2 > gpurun_out/r2_config3a_8gpu.json 2> gpurun_out/r2_config3a_8gpu.err
$TR --nproc-per-node 8 --master-port 29703 bench.py --workload config3b --steps 5 --warmup 2 > gpurun_out/r2_config3b_8gpu.json 2> gpurun_out/r2_config3b_8gpu.err
$TR --nproc-per-node 4 --master-port 29704 bench.py --workload config3a --steps 5 --warmup 2 > gpurun_out/r2_config3a_4gpu.json 2>> gpurun_out/r2_config3a_8gpu.err
$TR --nproc-per-node 2 --master-port 29705 bench.py --workload config3a --steps 5 --warmup 2 > gpurun_out/r2_config3a_2gpu.json 2>> gpurun_out/r2_config3a_8gpu.err
for n in 2 4 8; do $TR --nproc-per-node $n --master-port 2971$n tests/dist_split_micrograph.py > gpurun_out/r2_dist_parity_${n}gpu.log 2>&1; tail -2 gpurun_out/r2_dist_parity_${n}gpu.log; done
