set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_2gpu.log; tail -6 gpurun_out/r02_pytest_gpu_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_s2_2gpu.json 2> gpurun_out/r02_bench_s2_2gpu.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_s2_2gpu.json | cut -c1-6000; tail -5 gpurun_out/r02_bench_s2_2gpu.err
