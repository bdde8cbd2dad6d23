set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2; df -h /dev/shm /tmp | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1.log; tail -5 gpurun_out/r02_pytest_gpu_1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_s2_a.json 2> gpurun_out/r02_bench_s2_a.err; tail -c 3000 gpurun_out/r02_bench_s2_a.json; tail -5 gpurun_out/r02_bench_s2_a.err
B="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-shapes --skip-roofline-probe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_s2_geno8 $B > gpurun_out/ncu_s2.log 2>&1
ncu -i gpurun_out/r02_s2_geno8.ncu-rep --page raw --csv > gpurun_out/r02_k_geno8_s2_ncu_full.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_geno8 --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_s3_geno8 $B --workload s3 > gpurun_out/ncu_s3.log 2>&1
ncu -i gpurun_out/r02_s3_geno8.ncu-rep --page raw --csv > gpurun_out/r02_k_geno8_s3_ncu_full.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_geno|k_fq" -c 60 --csv --log-file gpurun_out/r02_launches_s2.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 200 ./vargeno_b200/tools/probes/sector_probe 32 0 sweep > gpurun_out/r02_sector_probe_sweep.jsonl 2>&1
tail -3 gpurun_out/r02_sector_probe_sweep.jsonl
ls -la gpurun_out | head -30
