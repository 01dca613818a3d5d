cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/r2_gpu_tests_multi_n8.log 2>&1; tail -3 $O/r2_gpu_tests_multi_n8.log
timeout 300 python tools/prof_scan.py 3 sharded 2>/dev/null | grep '^{' | tail -1 > $O/scan_sharded_r2.json; cat $O/scan_sharded_r2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 2> $O/bench_r2_n8.err | grep '^{' > $O/bench_r2_n8.json; python -c "
import json; d=json.load(open('$O/bench_r2_n8.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('per_step'))"
