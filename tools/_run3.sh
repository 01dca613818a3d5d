cd /root/repo
O=gpurun_out
timeout 400 ncu --clock-control none --set full --import-source on --kernel-name regex:k_scan_match --launch-skip 2 --launch-count 1 -o $O/r2b_scan_match -f python tools/prof_scan.py 3 > $O/r2b_scan_match.log 2>&1
ls -la $O/r2b_scan_match.ncu-rep
