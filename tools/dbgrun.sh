# debugging aid: SM cycles of every tensor-core convolution / weight-gradient launch of one calibration step (NQ_TC_DBG),
# optional env overrides as arguments ("base" = none)
run() { tag=$1; shift; env "$@" NQ_GRAPH=0 NQ_TC_DBG=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-hadamard-record --decode-steps 1 > gpurun_out/dbg_$tag.json 2> gpurun_out/dbg_$tag.err; echo "== $tag"; grep NQ_TC_DBG gpurun_out/dbg_$tag.err | tail -21 | grep -E "hw=(320x640|160x320|40x80)"; }
for t in "$@"; do
  case $t in
    base) run base X=1;;
    *) run "$t" "$t";;
  esac
done
