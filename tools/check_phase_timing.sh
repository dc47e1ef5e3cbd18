#!/bin/bash
# Runs the TIMING instantiations of the scan kernel (TA_PHASE_TIMING=1) on both staging paths and the one-hot pair path:
# one short pass each; prints the phase shares or the library error.
run() { out=$(env "$@" TA_PHASE_TIMING=1 timeout 60 python tools/profile_scan.py --config ${CFG:-C2} --passes 1 2>&1); if echo "$out" | grep -q "rror"; then echo "FAIL: $*"; echo "$out" | grep "NativeError" | cut -c1-300; else echo "ok:   $* | $(echo "$out" | grep 'phase cycles' | cut -c40-260)"; fi; }
run TA_X=1
run TA_NO_TMA=1
run TA_PAIR_PATH=onehot
