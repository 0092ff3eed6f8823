# The round's profile captures (one GPU): the launch list of the bench command with DRAM bytes per launch, then
# `ncu --set full` of the tensor-core convolution and weight-gradient launches of one calibration step.
#   bash tools/capture_round.sh <tag>      ->  gpurun_out/<tag>_launches.csv, <tag>_conv.ncu-rep, <tag>_wgrad.ncu-rep
set -e
TAG=$1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hadamard-record --decode-steps 1"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
export NQ_GRAPH=0
$CMD > gpurun_out/${TAG}_plain_eager.json 2> gpurun_out/${TAG}_plain_eager.err
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel|head_" -s 75 -c 25 \
    -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}*
