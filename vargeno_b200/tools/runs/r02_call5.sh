set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bgzf.py tests/test_gpu_cli.py -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_5.log; tail -25 gpurun_out/r02_pytest_gpu_5.log
timeout 900 python -m vargeno_b200.tools.cli_e2e --reads 64000000 --gpus 1 > gpurun_out/r02_cli_e2e_1gpu.jsonl 2> gpurun_out/r02_cli_e2e_1gpu.err; cat gpurun_out/r02_cli_e2e_1gpu.jsonl; tail -5 gpurun_out/r02_cli_e2e_1gpu.err
