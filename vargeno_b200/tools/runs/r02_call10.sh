set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bgzf.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_10.log; tail -6 gpurun_out/r02_pytest_gpu_10.log
timeout 600 python -m vargeno_b200.tools.inflate_bench --reads 6000000 --chunk-mb 512 > gpurun_out/r02_inflate_bench_v4.jsonl 2> gpurun_out/r02_inflate_bench_v4.err; cat gpurun_out/r02_inflate_bench_v4.jsonl; tail -3 gpurun_out/r02_inflate_bench_v4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_inflate_v4 python -m vargeno_b200.tools.inflate_bench --reads 2000000 --chunk-mb 512 --levels 1 --repeats 1 > gpurun_out/ncu_inflate.log 2>&1
ncu -i gpurun_out/r02_inflate_v4.ncu-rep --page raw --csv > gpurun_out/r02_k_inflate_v4_ncu_full.csv 2>/dev/null
ncu -i gpurun_out/r02_inflate_v4.ncu-rep --page source --csv > gpurun_out/r02_k_inflate_v4_source.csv 2>/dev/null
