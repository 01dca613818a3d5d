cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring or window or full_shape or reduce or positional" > $O/r2b_ring_tests.log 2>&1; tail -3 $O/r2b_ring_tests.log
for f in counter util; do PROF_FIELDS=$f python tools/prof_ring.py 5 survey 2>&1 | tail -1; done
for s in survey uniform walk mw; do python tools/prof_ring.py 5 $s 2>&1 | tail -1; done
