#!/bin/bash
# compute-sanitizer over one parity test that runs k_conv_tcr (the default conv path): memcheck, then racecheck (summary only)
mkdir -p gpurun_out
T='tests/test_gpu_parity.py::test_conv_kernel_paths_match_reference_golden[2]'
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_tcr.log python -m pytest "$T" -m gpu -q -x --timeout 800 --timeout-method thread > gpurun_out/sanitizer_memcheck_tcr.out 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_tcr.log; tail -2 gpurun_out/sanitizer_memcheck_tcr.out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/sanitizer_racecheck_tcr.log python -m pytest "$T" -m gpu -q -x --timeout 1100 --timeout-method thread > gpurun_out/sanitizer_racecheck_tcr.out 2>&1
echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/sanitizer_racecheck_tcr.log; tail -3 gpurun_out/sanitizer_racecheck_tcr.log; tail -2 gpurun_out/sanitizer_racecheck_tcr.out
grep "Race reported" gpurun_out/sanitizer_racecheck_tcr.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -30 > gpurun_out/sanitizer_racecheck_tcr_summary.txt
