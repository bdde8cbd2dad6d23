set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bench.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_20.log; tail -25 gpurun_out/r02_pytest_gpu_20.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_s2.json 2> gpurun_out/r02_final_bench_s2.err; echo "bench rc=$?"; python -c "
import json
for l in open('gpurun_out/r02_final_bench_s2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['frac'], d['e2e']['value']); print(json.dumps(d['cpu_baseline'])[:1500])
"; tail -3 gpurun_out/r02_final_bench_s2.err
