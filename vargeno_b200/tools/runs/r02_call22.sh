set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python bench.py > gpurun_out/r02_final_bench_s2.json 2> gpurun_out/r02_final_bench_s2.err ) 2>&1 | tail -3; python -c "
import json
for l in open('gpurun_out/r02_final_bench_s2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['steps'], d['warmup'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value']); print(json.dumps(d['cpu_baseline'])[:2500])
"; tail -3 gpurun_out/r02_final_bench_s2.err
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
