cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2b_gpu_tests_full.log 2>&1; tail -3 $O/r2b_gpu_tests_full.log
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:round(d[k],4) for k in ("filter_ms","prefix_ms","match_ms","phases_sum_ms","device_total_ms","hits")})'
timeout 200 python tools/prof_scan.py 6 2>/dev/null | python -c "$P"
