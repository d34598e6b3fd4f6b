#!/bin/bash
# First gpurun call of round 2 (one B200, ~6 GPU-minutes): (1) the GPU tests written after round 1's GPU minutes were spent,
# (2) the whole GPU suite, (3) the reference arm and the default bench line back to back, as the driver runs them.
#   gpurun --timeout 900 -- 'bash profiles/scripts/first_call_r02.sh'
set -x
O=gpurun_out
T=r02a
LTXB200_UNVERIFIED_TESTS=1 python -m pytest tests/test_zz_unverified_gpu.py -m gpu -q -s > $O/${T}_unverified_tests.log 2>&1
tail -15 $O/${T}_unverified_tests.log
python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1
tail -3 $O/${T}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_final.json 2> $O/${T}_bench_final.err
cat $O/${T}_bench_reference.json $O/${T}_bench_final.json | cut -c1-600
