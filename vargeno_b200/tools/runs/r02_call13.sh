set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shaped.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py tests/test_gpu_bgzf.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_13.log; tail -4 gpurun_out/r02_pytest_gpu_13.log
VGB200_LIB=$PWD/vargeno_b200/libvgb200_old.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" --tag old > gpurun_out/r02_ab2_old.jsonl 2>&1; tail -1 gpurun_out/r02_ab2_old.jsonl | cut -c1-330
timeout 500 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 "" VGB_NO_TAIL_OVERLAP=1 --tag staged > gpurun_out/r02_ab2_staged.jsonl 2>&1; cat gpurun_out/r02_ab2_staged.jsonl | cut -c1-330
