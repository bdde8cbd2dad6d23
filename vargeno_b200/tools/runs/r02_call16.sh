set -x
cd $GRAFT_REPO_ROOT
for v in "" _infA _infB; do
  VGB200_LIB=$PWD/vargeno_b200/libvgb200$v.so timeout 400 python -m vargeno_b200.tools.inflate_bench --reads 6000000 --chunk-mb 512 > gpurun_out/r02_inflate_var$v.jsonl 2> gpurun_out/r02_inflate_var$v.err; cut -c1-330 gpurun_out/r02_inflate_var$v.jsonl | sed 's/"text_bytes.*"inflate_plus/"inflate_plus/'; tail -2 gpurun_out/r02_inflate_var$v.err
done
