set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_12.log; tail -4 gpurun_out/r02_pytest_gpu_12.log
timeout 500 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 "" VGB_NO_TAIL_OVERLAP=1 --tag overlap > gpurun_out/r02_ab_overlap.jsonl 2>&1; cat gpurun_out/r02_ab_overlap.jsonl | cut -c1-330
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s2_geno8 $B > gpurun_out/ncu_s2f.log 2>&1
ncu -i gpurun_out/r02f_s2_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s2_ncu_full.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s3_geno8 $B --workload s3 > gpurun_out/ncu_s3f.log 2>&1
ncu -i gpurun_out/r02f_s3_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s3_ncu_full.csv 2>/dev/null
python -m vargeno_b200.tools.ncu_constants s2=gpurun_out/r02_final_k_geno8_s2_ncu_full.csv s3=gpurun_out/r02_final_k_geno8_s3_ncu_full.csv > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2.json 2> gpurun_out/r02_final_bench_s2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_final_bench_s2.json; tail -3 gpurun_out/r02_final_bench_s2.err
timeout 600 python -m vargeno_b200.tools.cli_e2e --reads 64000000 --gpus 1 --skip-gzip > gpurun_out/r02_final_cli_e2e_1gpu.jsonl 2> gpurun_out/r02_final_cli_e2e_1gpu.err; cat gpurun_out/r02_final_cli_e2e_1gpu.jsonl
