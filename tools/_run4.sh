cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gpm or poller" > $O/r2b_gpm_tests.log 2>&1; tail -15 $O/r2b_gpm_tests.log
