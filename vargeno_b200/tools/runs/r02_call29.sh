set -x
cd $GRAFT_REPO_ROOT
for v in _old ""; do
  VGB200_LIB=$PWD/vargeno_b200/libvgb200$v.so timeout 300 python bench.py --workload s1 --skip-cpu --skip-roofline-probe --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('S1 lib$v', 'value %.1f M' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'k_geno %.4f' % d['kernel_ms_per_step']['k_geno'], 'serial %.1f M' % (d['roofline']['serial_reads_per_s']/1e6))
"
done
VGB200_LIB=$PWD/vargeno_b200/libvgb200_old.so timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 --tag old > gpurun_out/r02_ab7_old.jsonl 2>&1; grep variant gpurun_out/r02_ab7_old.jsonl | cut -c1-300
timeout 400 python -m vargeno_b200.tools.sweep_wgs "" VGB_NO_TAIL_OVERLAP=1 --tag new > gpurun_out/r02_ab7_new.jsonl 2>&1; grep variant gpurun_out/r02_ab7_new.jsonl | cut -c1-300
