set -e
python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
