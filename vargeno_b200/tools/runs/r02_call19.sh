set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_19.log; tail -3 gpurun_out/r02_pytest_gpu_19.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_final_bench_ref_s2.json 2> gpurun_out/r02_final_bench_ref_s2.err ) 2>&1 | tail -4; echo "ref rc=$?"; cut -c1-1500 gpurun_out/r02_final_bench_ref_s2.json; tail -3 gpurun_out/r02_final_bench_ref_s2.err
