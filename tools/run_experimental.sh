#!/bin/bash
# One GPU call for everything that was written after round 1's GPU budget was spent ("Next"):
#   gpurun --timeout 300 -- 'bash tools/run_experimental.sh'
# 1. the GPU golden test of the fused chains (reference-generated fixture)
# 2. the experimental paths' parity tests (Legendre run table, L2-prefetch variant of the white A-matvec)
# 3. their timings against the default paths (tools/fused_probe.py with CM2_EXPERIMENTAL=1)
# Outputs go to gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_zz_fused_golden.py -q > gpurun_out/exp_golden.log 2>&1
echo "golden rc=$?"; tail -3 gpurun_out/exp_golden.log
CM2_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_fused_chains.py -q -k "poly_run_table or l2_prefetch" > gpurun_out/exp_tests.log 2>&1
echo "experimental tests rc=$?"; tail -5 gpurun_out/exp_tests.log
CM2_EXPERIMENTAL=1 timeout 200 python tools/fused_probe.py > gpurun_out/exp_probe.json 2> gpurun_out/exp_probe.err
echo "probe rc=$?"; cat gpurun_out/exp_probe.json; tail -3 gpurun_out/exp_probe.err
