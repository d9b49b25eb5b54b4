#!/bin/bash
# Per-source-line instruction counts of the FIRST match of a C3 registration (far queries).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/one_reg.py <<'PY'
import sys
sys.path.insert(0, ".")
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair
X_fix, X_mov, _ = make_pair(1_000_000, 0)
with _capi.Engine() as e:
    r = sb.register(X_fix, X_mov, correspondences=100_000, engine=e, want_normals=False)
    print(r.iterations)
PY
ncu --set full --import-source on --clock-control none -k "regex:k_match_grid_coop" -c 1 -f -o gpurun_out/first_match python /tmp/one_reg.py > gpurun_out/first_match.log 2>&1
true
ncu -i gpurun_out/first_match.ncu-rep --page raw --csv > gpurun_out/first_match_raw.csv 2>> gpurun_out/first_match.log
ls -la gpurun_out/first_match* ; head -c 600 gpurun_out/first_match_source.csv
