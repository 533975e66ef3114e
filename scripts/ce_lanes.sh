#!/bin/bash
for L in 1 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$L bench.py --gpus 2 --steps 20 --warmup 3 --power-iters 10 --ce-lanes $L --no-fused > gpurun_out/ce_l$L.json 2> gpurun_out/ce_l$L.err
echo "lanes $L"; grep "^\[C" gpurun_out/ce_l$L.err | sort | uniq | grep copy; grep -i "Traceback\|Error" gpurun_out/ce_l$L.err | head -3
done
