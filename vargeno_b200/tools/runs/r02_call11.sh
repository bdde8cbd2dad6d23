set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shaped.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu_11.log; tail -3 gpurun_out/r02_pytest_gpu_11.log
VGB200_LIB=$PWD/vargeno_b200/libvgb200_old.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" "" --tag old > gpurun_out/r02_ab_old1.jsonl 2>&1; tail -2 gpurun_out/r02_ab_old1.jsonl | cut -c1-330
timeout 400 python -m vargeno_b200.tools.sweep_wgs "" "" --tag new > gpurun_out/r02_ab_new1.jsonl 2>&1; tail -2 gpurun_out/r02_ab_new1.jsonl | cut -c1-330
VGB200_LIB=$PWD/vargeno_b200/libvgb200_old.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" --tag old > gpurun_out/r02_ab_old2.jsonl 2>&1; tail -1 gpurun_out/r02_ab_old2.jsonl | cut -c1-330
timeout 400 python -m vargeno_b200.tools.sweep_wgs "" --tag new > gpurun_out/r02_ab_new2.jsonl 2>&1; tail -1 gpurun_out/r02_ab_new2.jsonl | cut -c1-330
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s2_geno8 $B > gpurun_out/ncu_s2f.log 2>&1
ncu -i gpurun_out/r02f_s2_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s2_ncu_full.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s3_geno8 $B --workload s3 > gpurun_out/ncu_s3f.log 2>&1
ncu -i gpurun_out/r02f_s3_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s3_ncu_full.csv 2>/dev/null
ls -la gpurun_out/*final_k_geno8*
