cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29614 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n8.json') if l.startswith('{')][-1]); print(d['value'], json.dumps(d.get('ulysses'))[:900])"
timeout 400 $TR --master-port 29611 tools/ulysses_check.py --frames 21 --graph 1 --timeline 1 > gpurun_out/uly_P8.log 2>&1; echo rc=$?; tail -1 gpurun_out/uly_P8.log | cut -c1-3000
