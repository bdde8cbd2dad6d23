set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shaped.py tests/test_gpu_fullsize.py tests/test_gpu_bench.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_25.log; tail -4 gpurun_out/r02_pytest_gpu_25.log
VGB200_LIB=$PWD/vargeno_b200/libvgb200_old.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 --tag old > gpurun_out/r02_ab5_old.jsonl 2>&1; grep variant gpurun_out/r02_ab5_old.jsonl | cut -c1-330
timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 --tag bskip > gpurun_out/r02_ab5_new.jsonl 2>&1; grep variant gpurun_out/r02_ab5_new.jsonl | cut -c1-330
