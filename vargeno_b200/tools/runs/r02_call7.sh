set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_7.log; tail -6 gpurun_out/r02_pytest_gpu_7.log
timeout 600 python -m vargeno_b200.tools.inflate_bench --reads 6000000 --chunk-mb 512 > gpurun_out/r02_inflate_bench.jsonl 2> gpurun_out/r02_inflate_bench.err; cat gpurun_out/r02_inflate_bench.jsonl; tail -3 gpurun_out/r02_inflate_bench.err
timeout 600 python -m vargeno_b200.tools.inflate_bench --reads 6000000 --chunk-mb 128 --levels 6 >> gpurun_out/r02_inflate_bench.jsonl 2>> gpurun_out/r02_inflate_bench.err; tail -1 gpurun_out/r02_inflate_bench.jsonl
timeout 600 python -m vargeno_b200.tools.cli_e2e --reads 64000000 --gpus 1 --skip-gzip > gpurun_out/r02d_cli_e2e_1gpu.jsonl 2> gpurun_out/r02d_cli_e2e_1gpu.err; cat gpurun_out/r02d_cli_e2e_1gpu.jsonl; tail -5 gpurun_out/r02d_cli_e2e_1gpu.err
