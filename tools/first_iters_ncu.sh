#!/bin/bash
# Durations (ncu, serialised, cold caches) of the match / reject+solve launches of two C3 registrations:
# shows what the first two iterations cost next to the steady state.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/two_regs.py <<'PY'
import sys
sys.path.insert(0, ".")
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair
X_fix, X_mov, _ = make_pair(1_000_000, 0)
with _capi.Engine() as e:
    for i in range(2):
        r = sb.register(X_fix, X_mov, correspondences=100_000, engine=e, want_normals=False)
    print(r.iterations)
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --kernel-name regex:"k_match_grid_coop|k_reject_solve|k_rs_fused" \
    --csv --log-file gpurun_out/first_iters.csv python /tmp/two_regs.py > gpurun_out/first_iters.log 2>&1
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open("gpurun_out/first_iters.csv") if not l.startswith("==")]))
by = {}
for r in rows:
    by.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0][-24:]})[r["Metric Name"]] = r["Metric Value"]
for k, v in by.items():
    print(k, v["name"], v.get("gpu__time_duration.sum"), v.get("smsp__inst_executed.sum"))
PY
