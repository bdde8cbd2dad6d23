set -x
cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2_${N}gpu.json 2> gpurun_out/r02_final_bench_s2_${N}gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_final_bench_s2_${N}gpu.json; tail -3 gpurun_out/r02_final_bench_s2_${N}gpu.err
