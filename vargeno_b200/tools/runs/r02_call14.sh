set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2.json 2> gpurun_out/r02_final_bench_s2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_final_bench_s2.json; tail -3 gpurun_out/r02_final_bench_s2.err
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 72 --csv --log-file gpurun_out/r02_final_launches_s2.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 72 --csv --log-file gpurun_out/r02_final_launches_s3.csv $B --workload s3 > gpurun_out/ncu_ll3.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --workload s1 > gpurun_out/r02_final_bench_s1.json 2> gpurun_out/r02_final_bench_s1.err; cut -c1-300 gpurun_out/r02_final_bench_s1.json
timeout 900 python -m vargeno_b200.tools.cli_e2e --reads 128000000 --gpus 1 > gpurun_out/r02_final_cli_e2e_1gpu.jsonl 2> gpurun_out/r02_final_cli_e2e_1gpu.err; cat gpurun_out/r02_final_cli_e2e_1gpu.jsonl
