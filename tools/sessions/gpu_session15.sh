# round 2, session 15: compute-sanitizer (memcheck, racecheck, synccheck) over tools/sanitize_target.py
mkdir -p gpurun_out
which compute-sanitizer; compute-sanitizer --version | head -3
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python tools/sanitize_target.py > gpurun_out/s15_sanitizer_$tool.log 2>&1
  echo "rc=$?"; tail -5 gpurun_out/s15_sanitizer_$tool.log
done
