set -x
cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_6.log; tail -5 gpurun_out/r02_pytest_gpu_6.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_s2_c.json 2> gpurun_out/r02_bench_s2_c.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02_bench_s2_c.json; tail -3 gpurun_out/r02_bench_s2_c.err
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 60 --csv --log-file gpurun_out/r02c_launches_s2.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 60 --csv --log-file gpurun_out/r02c_launches_s3.csv $B --workload s3 > gpurun_out/ncu_ll3.log 2>&1
timeout 600 python -m vargeno_b200.tools.cli_e2e --reads 64000000 --gpus 1 > gpurun_out/r02c_cli_e2e_1gpu.jsonl 2> gpurun_out/r02c_cli_e2e_1gpu.err; cat gpurun_out/r02c_cli_e2e_1gpu.jsonl; tail -5 gpurun_out/r02c_cli_e2e_1gpu.err
