set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_2.log; tail -5 gpurun_out/r02_pytest_gpu_2.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/r02_bench_s2_b.json 2> gpurun_out/r02_bench_s2_b.err; tail -c 2500 gpurun_out/r02_bench_s2_b.json; tail -5 gpurun_out/r02_bench_s2_b.err
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --workload s1 > gpurun_out/r02_bench_s1_b.json 2> gpurun_out/r02_bench_s1_b.err; head -c 1500 gpurun_out/r02_bench_s1_b.json; tail -5 gpurun_out/r02_bench_s1_b.err
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02b_s2_geno8 $B > gpurun_out/ncu_s2b.log 2>&1
ncu -i gpurun_out/r02b_s2_geno8.ncu-rep --page raw --csv > gpurun_out/r02b_k_geno8_s2_ncu_full.csv 2>/dev/null
ncu -i gpurun_out/r02b_s2_geno8.ncu-rep --page source --csv > gpurun_out/r02b_s2_sass.csv 2>/dev/null
nvidia-smi --query-gpu=memory.used --format=csv
