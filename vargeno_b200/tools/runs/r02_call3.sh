set -x
cd $GRAFT_REPO_ROOT
timeout 300 ./vargeno_b200/tools/probes/sector_probe 32 0 sweep > gpurun_out/r02_sector_probe_sweep2.jsonl 2>&1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k probe 2>&1 | tail -3
timeout 600 python -m vargeno_b200.tools.sweep_wgs "" VGB_CARVEOUT=25 VGB_CARVEOUT=50 VGB_CARVEOUT=75 VGB_CARVEOUT=100 VGB_GENO4_MINB=3 VGB_GENO4_MINB=5 "" --probe > gpurun_out/r02_sweep_wgs.jsonl 2>&1
VGB200_LIB=$PWD/vargeno_b200/libvgb200_ev16.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_CARVEOUT=50 --tag ev16 > gpurun_out/r02_sweep_wgs_ev16.jsonl 2>&1
VGB200_LIB=$PWD/vargeno_b200/libvgb200_ev12.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_CARVEOUT=50 --tag ev12 > gpurun_out/r02_sweep_wgs_ev12.jsonl 2>&1
cat gpurun_out/r02_sweep_wgs*.jsonl
free -g | head -2
( timeout 1700 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_s2.json 2> gpurun_out/r02_bench_ref_s2.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_ref_s2.json; tail -5 gpurun_out/r02_bench_ref_s2.err ) 
