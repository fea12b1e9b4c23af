# A/B harness for the sketch kernel on the GPU box: rebuilds mm_sketch.cu with different -D flags and times K1 alone
set -x
cd $GRAFT_REPO_ROOT
python scripts/k1_perf.py 200000 > gpurun_out/k1_base.log 2>&1
for v in "$@"; do
  touch mashmap_b200/csrc/mm_sketch.cu
  make -C mashmap_b200/csrc EXTRA="$v" > /dev/null 2>&1
  tag=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  grep -A3 "8k_sketchILi19E" mashmap_b200/csrc/build/mm_sketch.ptxas.log | grep -E "registers|spill" > gpurun_out/k1_$tag.log
  python scripts/k1_perf.py 200000 >> gpurun_out/k1_$tag.log 2>&1
done
tail -n 2 gpurun_out/k1_*.log
