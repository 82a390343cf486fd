cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 tools/ulysses_check.py --frames 21 --graph 1 --timeline 1 > gpurun_out/uly_P4.log 2>&1; echo rc=$?; tail -1 gpurun_out/uly_P4.log | cut -c1-1500
LLB_PDL=1 timeout 400 $TR --master-port 29612 tools/ulysses_check.py --frames 21 --graph 1 --out gpurun_out/ulysses_P4_graph1_pdl.json > gpurun_out/uly_P4_pdl.log 2>&1; echo rc=$?; tail -1 gpurun_out/uly_P4_pdl.log | cut -c1-700
