#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_forward.py > gpurun_out/sanitize_memcheck_forward_r02.log 2>&1
grep -E "forwards ok|IRR_PWC|PWCNet|ERROR SUMMARY|Invalid|Illegal" gpurun_out/sanitize_memcheck_forward_r02.log | head -12
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke_r02.log 2>&1
grep -E "smoke ok|ERROR SUMMARY|Invalid|Illegal" gpurun_out/sanitize_memcheck_smoke_r02.log | head -5
