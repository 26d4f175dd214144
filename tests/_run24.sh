mkdir -p gpurun_out/s4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 50 --warmup 5 2> gpurun_out/s4/bench8.err | tail -c 1300
