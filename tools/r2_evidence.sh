#!/bin/bash
# Round-2 evidence run on a B200 box (through gpurun): GPU tests, the bench and reference arms, the ncu launch list of the bench
# command and one `ncu --set full` capture of each dominant kernel.  Everything lands in gpurun_out/ (summaries: tools/ncu_summary.py).
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2_gpu_tests_full.log 2>&1; tail -3 $O/r2_gpu_tests_full.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> $O/bench_r2_ref.err | grep '^{' > $O/bench_r2_ref.json
timeout 900 python bench.py 2> $O/bench_r2_n1.err | grep '^{' > $O/bench_r2_n1.json; tail -c 300 $O/bench_r2_n1.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name regex:"k_" -c 400 --csv --log-file $O/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-shapes --no-verify > $O/b_under_ncu_r2.log 2>&1
FULL="$NCU --set full --import-source on --kernel-name-base demangled"
timeout 400 $FULL --kernel-name regex:"k_window_reduce<.bool.1, .int.15, .bool.0" --launch-skip 3 --launch-count 1 -o $O/r2_win_survey -f python tools/prof_ring.py 5 survey > $O/r2_win_survey.log 2>&1
timeout 400 $FULL --kernel-name regex:"k_window_reduce<.bool.1, .int.16, .bool.1" --launch-skip 2 --launch-count 1 -o $O/r2_range_pass -f python tools/prof_ring.py 2 survey range > $O/r2_range_pass.log 2>&1
timeout 400 $FULL --kernel-name regex:k_scan_filter --launch-skip 2 --launch-count 1 -o $O/r2_scan_filter -f python tools/prof_scan.py 3 > $O/r2_scan_filter.log 2>&1
timeout 400 $FULL --kernel-name regex:k_scan_match --launch-skip 2 --launch-count 1 -o $O/r2_scan_match -f python tools/prof_scan.py 3 > $O/r2_scan_match.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --kernel-name regex:"k_scan|k_cand" -c 30 --csv --log-file $O/launches_scan_r2.csv python tools/prof_scan.py 3 > /dev/null 2>&1
timeout 200 python tools/prof_scan.py 6 cpu 2>/dev/null | grep '^{' > $O/scan_r2.json
for s in uniform mw temp const walk; do python tools/prof_ring.py 5 $s 2>&1 | tail -1; done > $O/r2_shapes.txt
for f in counter temp power util; do PROF_FIELDS=$f python tools/prof_ring.py 5 survey 2>&1 | tail -1; done >> $O/r2_shapes.txt
python tools/exp_scan_density.py 2>&1 | tail -10 > $O/r2_scan_by_hit_kind.txt
ls -la $O/*.ncu-rep | tail -5
