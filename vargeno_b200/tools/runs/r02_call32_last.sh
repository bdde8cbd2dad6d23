set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err; echo "bench rc=$?"; python -c "
import json
for l in open('gpurun_out/r02_last_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['parity']['parity_checked'], d['gpu_launches'], d['clocks'])
"
