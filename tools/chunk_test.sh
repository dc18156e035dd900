for c in 4096 8192 16384 32768; do python tools/quick_perf.py 4096 8 10 1048576 $c 2>&1 | tail -1; done
for c in 8192 32768 65536; do python tools/quick_perf.py 1024 6 10 1000000 $c 2>&1 | tail -1; done
