#!/bin/bash
# compute-sanitizer passes over the hot path at small sizes: memcheck of the whole IRR_PWC smoke forward, racecheck of the
# correlation kernels (mbarrier-ordered shared-memory rings) — summaries go to profiles/.
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1
grep -E "smoke ok|ERROR SUMMARY|Invalid|Illegal" gpurun_out/sanitize_memcheck_smoke.log | head -5
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python scripts/corr_one.py 2 32 40 64 fused > gpurun_out/sanitize_racecheck_corr_fused.log 2>&1
grep -E "max \||RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck_corr_fused.log | head -5
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python scripts/corr_one.py 2 32 40 64 > gpurun_out/sanitize_racecheck_corr_plain.log 2>&1
grep -E "max \||RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck_corr_plain.log | head -5
