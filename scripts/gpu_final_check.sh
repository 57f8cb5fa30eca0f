cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -n 2 gpurun_out/pytest_gpu.log | tr '\n' ' ')"
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke(); g.smoke_stream()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) kernels-with-hazards: $(grep -o 'at void tapir::[^(]*' gpurun_out/sanitizer_$tool.log | sort -u | tr '\n' ' ')"
done
