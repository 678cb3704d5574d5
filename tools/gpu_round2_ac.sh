#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "wide_rates or argument" 2>&1 | tail -5
timeout 200 python tests/fuzz_parity.py --seconds 100 --seed 23 > $O/r2_fuzz_ac.log 2>&1; tail -2 $O/r2_fuzz_ac.log; grep -c 'fe=tc' $O/r2_fuzz_ac.log; grep -c 'D=24\|D=32' $O/r2_fuzz_ac.log
