for z in 0 1 2; do WN_EXP_Z=$z python tools/host_overhead.py | tail -1; done
