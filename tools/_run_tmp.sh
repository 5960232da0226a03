bash tools/gpu_sanitize.sh r2 2>&1 | tee gpurun_out/sanitizer_r2_summary.txt
