mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 profiles/tools/timeline.py --tag n2 > gpurun_out/r2e_tl2.log 2>&1; echo "tl2 rc=$?"
B200GS_PEER_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 profiles/tools/timeline.py --tag n2nccl > gpurun_out/r2e_tl2n.log 2>&1; echo "tl2n rc=$?"
timeout 300 python profiles/tools/timeline.py --config 3 --mode gsplat --tag n1c3 > gpurun_out/r2e_tl1.log 2>&1; echo "tl1 rc=$?"
tail -3 gpurun_out/r2e_tl2.log gpurun_out/r2e_tl1.log
