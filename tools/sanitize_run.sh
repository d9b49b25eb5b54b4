#!/bin/bash
# run inside gpurun: the probe under every sanitizer tool, bounded
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/sanitize_probe.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/san_plain.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --launch-timeout 120 python tools/sanitize_probe.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|hazard|Error" gpurun_out/san_$tool.log | head -8
done
