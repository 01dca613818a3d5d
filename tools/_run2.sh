cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "scan or kmsg" > $O/r2b_scan_tests.log 2>&1; tail -3 $O/r2b_scan_tests.log
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:round(d[k],4) for k in ("filter_ms","prefix_ms","match_ms","phases_sum_ms","device_total_ms","hits")})'
echo default; timeout 200 python tools/prof_scan.py 6 2>/dev/null | python -c "$P"
for v in "$@"; do echo "variant $v"; GPUD_B200_LIB=build/libgpud_$v.so timeout 200 python tools/prof_scan.py 6 2>/dev/null | python -c "$P"; done
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name regex:"k_scan|k_cand" -c 30 --csv --log-file $O/launches_scan_r2b.csv python tools/prof_scan.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/launches_scan_r2b.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); d=defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:40]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in d.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us')
PY
