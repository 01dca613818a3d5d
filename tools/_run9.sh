cd /root/repo
O=gpurun_out
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> $O/bench_r2_ref.err | grep '^{' > $O/bench_r2_ref.json
timeout 900 python bench.py 2> $O/bench_r2_n1.err | grep '^{' > $O/bench_r2_n1.json; tail -c 200 $O/bench_r2_n1.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name regex:"k_" -c 400 --csv --log-file $O/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-shapes --no-verify > $O/b_under_ncu_r2.log 2>&1
FULL="$NCU --set full --import-source on --kernel-name-base demangled"
timeout 400 $FULL --kernel-name regex:"k_window_reduce<.bool.1, .int.15, .bool.0" --launch-skip 3 --launch-count 1 -o $O/r2_win_survey -f python tools/prof_ring.py 5 survey > $O/r2_win_survey.log 2>&1
python -c "
import json; d=json.load(open('$O/bench_r2_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['scan']['device_ms'], d['range']['device_ms'], d['verify']['ok'])"
