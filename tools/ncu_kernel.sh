# ncu --set full with SASS-level stall sampling of the launches of one kernel family in a calibration step.
#   bash tools/ncu_kernel.sh <regex> <skip> <count> <out-name>
set -e
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hadamard-record --decode-steps 1"
NQ_GRAPH=0 $CMD > gpurun_out/ncu_plain.log 2>&1
NQ_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 $CMD > gpurun_out/ncu_run.log 2>&1
ls -la gpurun_out/$4.ncu-rep
