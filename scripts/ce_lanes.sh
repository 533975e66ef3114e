#!/bin/bash
# 2-GPU probe of the copy-engine exchange (lanes).  Every multi-rank command runs under its own `timeout`: a hung rank
# otherwise holds the whole box until gpurun's limit (round 2 lost 150 GPU-minutes to exactly that).
for L in 1 2 4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$L bench.py --gpus 2 --steps 20 --warmup 3 --power-iters 10 --ce-lanes $L --no-fused > gpurun_out/ce_l$L.json 2> gpurun_out/ce_l$L.err
echo "lanes $L"; grep "^\[C" gpurun_out/ce_l$L.err | sort | uniq | grep copy; grep -i "Traceback\|Error" gpurun_out/ce_l$L.err | head -3
done
