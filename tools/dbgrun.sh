# debugging aid: SM cycles of every tensor-core convolution launch of one calibration step (NQ_TC_DBG), optional env overrides
run() { tag=$1; shift; env "$@" NQ_GRAPH=0 NQ_TC_DBG=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-hadamard-record --decode-steps 1 > gpurun_out/dbg_$tag.json 2> gpurun_out/dbg_$tag.err; echo "== $tag"; grep NQ_TC_DBG gpurun_out/dbg_$tag.err | tail -13 | grep -A1 -E "hw=(320x640|160x320|40x80)" | grep -v "^--"; }
for t in "$@"; do
  case $t in
    base) run base X=1;;
    skipboth) run skipboth NQ_TC_SKIP=5;;
    *) run "$t" "$t";;
  esac
done
