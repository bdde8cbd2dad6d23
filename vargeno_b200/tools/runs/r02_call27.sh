set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for v in _old ""; do
  VGB200_LIB=$PWD/vargeno_b200/libvgb200$v.so timeout 300 python bench.py --workload s1 --skip-cpu --skip-roofline-probe --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lib$v', 'value %.1f M' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'serial %.1f M' % (d['roofline']['serial_reads_per_s']/1e6))
"
done
done
