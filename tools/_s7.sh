mkdir -p gpurun_out
VKX_BLEND=tc timeout 420 compute-sanitizer --tool initcheck --print-limit 100000 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_default_initcheck.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|sanitize pass ok" gpurun_out/r02_sanitizer_default_initcheck.log | tail -2
grep -A1 "Uninitialized __global__" gpurun_out/r02_sanitizer_default_initcheck.log | grep " at " | sed 's/+0x.*//; s/.*at //' | cut -c1-120 | sort | uniq -c | sort -rn | head -20
