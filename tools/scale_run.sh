# One box, N GPUs (torchrun, NCCL): the calibration bench line, the 1080p 12M decode sweep point (BASELINE configs[4]) and the
# Omega search (configs[3]); JSON lines under gpurun_out/<tag>_*.json
#   bash tools/scale_run.sh <N> <tag>
N=$1; TAG=$2
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
$RUN bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_calib_${N}gpu.json 2> gpurun_out/${TAG}_calib_${N}gpu.err
$RUN bench.py --gpus $N --mode decode --workload hnerv-1080p-12m --batch 2 --steps 60 --warmup 5 > gpurun_out/${TAG}_decode1080p_${N}gpu.json 2> gpurun_out/${TAG}_decode1080p_${N}gpu.err
$RUN bench.py --gpus $N --mode decode --workload hnerv-bunny-3m --batch 2 --steps 200 --warmup 5 > gpurun_out/${TAG}_decode3m_${N}gpu.json 2> gpurun_out/${TAG}_decode3m_${N}gpu.err
$RUN bench.py --gpus $N --mode omega --steps 1 --warmup 0 > gpurun_out/${TAG}_omega_${N}gpu.json 2> gpurun_out/${TAG}_omega_${N}gpu.err
for f in calib decode1080p decode3m omega; do echo "== $f"; tail -n 1 gpurun_out/${TAG}_${f}_${N}gpu.json | cut -c1-400; tail -n 2 gpurun_out/${TAG}_${f}_${N}gpu.err | cut -c1-300; done
