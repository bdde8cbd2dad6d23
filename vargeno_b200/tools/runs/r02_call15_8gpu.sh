set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; nproc; free -g | head -2; nvidia-smi topo -m | head -11
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2_8gpu.json 2> gpurun_out/r02_final_bench_s2_8gpu.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_final_bench_s2_8gpu.json; tail -4 gpurun_out/r02_final_bench_s2_8gpu.err
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -q -k two_gpus 2>&1 | tail -3 > gpurun_out/r02_pytest_two_gpus.log; cat gpurun_out/r02_pytest_two_gpus.log
timeout 900 python -m vargeno_b200.tools.cli_e2e --reads 128000000 --gpus 8 --skip-gzip > gpurun_out/r02_final_cli_e2e_8gpu.jsonl 2> gpurun_out/r02_final_cli_e2e_8gpu.err; cat gpurun_out/r02_final_cli_e2e_8gpu.jsonl; tail -3 gpurun_out/r02_final_cli_e2e_8gpu.err
