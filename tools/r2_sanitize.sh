#!/bin/bash
# compute-sanitizer over the GPU tests of the kernels this round touched (ring / range / positional windows, scan); logs into gpurun_out/
cd "$(dirname "$0")/.."
O=gpurun_out
SEL="ring_windows or wrap_and_odd or tie_heavy or positional or special_values or quantiles or reduce_range or scan_golden or scan_each or scan_synthetic or scan_raw or scan_edges or scan_ragged or scan_anchor or scan_ext_matchers or scan_multiline or component_objects"
timeout 1500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $O/r2_san_mem.log 2>&1; tail -4 $O/r2_san_mem.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $O/r2_san_race.log 2>&1; tail -4 $O/r2_san_race.log
