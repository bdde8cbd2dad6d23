set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shaped.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_8.log; tail -6 gpurun_out/r02_pytest_gpu_8.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_inflate python -m vargeno_b200.tools.inflate_bench --reads 2000000 --chunk-mb 512 --levels 1 --repeats 1 > gpurun_out/ncu_inflate.log 2>&1
ncu -i gpurun_out/r02_inflate.ncu-rep --page raw --csv > gpurun_out/r02_k_inflate_ncu_full.csv 2>/dev/null
ncu -i gpurun_out/r02_inflate.ncu-rep --page source --csv > gpurun_out/r02_k_inflate_source.csv 2>/dev/null
timeout 600 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_WIDE=1 "" VGB_NO_WIDE=1 > gpurun_out/r02_sweep_wide.jsonl 2>&1; cat gpurun_out/r02_sweep_wide.jsonl
