# Final 8-GPU record of the round: calibration bench line, decode sweep points, and the north-star job data parallel
# (global mini-batch 2 per GPU) through the CLI.   bash tools/scale_run8.sh <N> <tag>
N=$1; TAG=$2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_calib_${N}gpu.json 2> gpurun_out/${TAG}_calib_${N}gpu.err
$RUN bench.py --gpus $N --mode decode --workload hnerv-1080p-12m --batch 2 --steps 60 --warmup 5 > gpurun_out/${TAG}_decode1080p_${N}gpu.json 2> gpurun_out/${TAG}_decode1080p_${N}gpu.err
$RUN bench.py --gpus $N --mode decode --workload hnerv-bunny-3m --batch 2 --steps 200 --warmup 5 > gpurun_out/${TAG}_decode3m_${N}gpu.json 2> gpurun_out/${TAG}_decode3m_${N}gpu.err
for f in calib decode1080p decode3m; do echo "== $f"; tail -n 1 gpurun_out/${TAG}_${f}_${N}gpu.json | cut -c1-330; done
timeout 400 $RUN tools/full_run.py --arch hnerv --batch $((2 * N)) --fp-epochs 5 --out gpurun_out/${TAG}_full_run_hnerv_${N}gpu.json > gpurun_out/${TAG}_full_run_${N}gpu.log 2>&1
tail -n 3 gpurun_out/${TAG}_full_run_${N}gpu.log | cut -c1-1200
