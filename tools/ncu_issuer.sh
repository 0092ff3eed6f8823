# ncu --set full with SASS-level stall sampling of the stage-5 forward and data-gradient kernels (one launch each)
set -e
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hadamard-record --decode-steps 1"
NQ_GRAPH=0 $CMD > gpurun_out/ncu_plain.log 2>&1
NQ_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 28 -c 12 -o gpurun_out/r02n_conv $CMD > gpurun_out/ncu_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
