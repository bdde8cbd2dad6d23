set -x
cd $GRAFT_REPO_ROOT
timeout 500 python -m vargeno_b200.tools.sweep_wgs "" VGB_FQ_VARIANT=3 VGB_FQ_VARIANT=4 VGB_FQ_VARIANT=5 VGB_FQ_VARIANT=6 "" VGB_FQ_VARIANT=3 --tag framing > gpurun_out/r02_ab_framing.jsonl 2>&1; grep variant gpurun_out/r02_ab_framing.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-22s framing %.4f ms  geno %.4f  reads/s %.1f M | s3 %.1f M' % (d['variant'], d['s2']['framing_ms'], d['s2']['k_geno_ms'], d['s2']['reads_per_s']/1e6, d['s3']['reads_per_s']/1e6))
"
VGB_FQ_VARIANT=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_bgzf.py tests/test_gpu_fullsize.py tests/test_gpu_bench.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu_21.log; tail -4 gpurun_out/r02_pytest_gpu_21.log
