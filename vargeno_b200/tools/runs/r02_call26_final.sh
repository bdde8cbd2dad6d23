set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.log; tail -3 gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s2_geno8 $B > gpurun_out/ncu_s2f.log 2>&1
ncu -i gpurun_out/r02f_s2_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s2_ncu_full.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02f_s3_geno8 $B --workload s3 > gpurun_out/ncu_s3f.log 2>&1
ncu -i gpurun_out/r02f_s3_geno8.ncu-rep --page raw --csv > gpurun_out/r02_final_k_geno8_s3_ncu_full.csv 2>/dev/null
cp gpurun_out/r02_final_k_geno8_s2_ncu_full.csv profiles/r02_k_geno8_s2_ncu_full.csv; cp gpurun_out/r02_final_k_geno8_s3_ncu_full.csv profiles/r02_k_geno8_s3_ncu_full.csv
python -m vargeno_b200.tools.ncu_constants s2=profiles/r02_k_geno8_s2_ncu_full.csv s3=profiles/r02_k_geno8_s3_ncu_full.csv > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 72 --csv --log-file gpurun_out/r02_final_launches_s2.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 72 --csv --log-file gpurun_out/r02_final_launches_s3.csv $B --workload s3 > gpurun_out/ncu_ll3.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2.json 2> gpurun_out/r02_final_bench_s2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_final_bench_s2.json; tail -3 gpurun_out/r02_final_bench_s2.err
timeout 900 python bench.py > gpurun_out/r02_final_bench_s2_default.json 2> gpurun_out/r02_final_bench_s2_default.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --workload s1 > gpurun_out/r02_final_bench_s1.json 2> gpurun_out/r02_final_bench_s1.err; cut -c1-200 gpurun_out/r02_final_bench_s1.json
