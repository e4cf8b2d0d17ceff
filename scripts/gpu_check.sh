#!/bin/bash
# One GPU-box call: the GPU test-suite, then device timings at the bench shapes (developer tool).
#   scripts/gpu_check.sh <tag> ["<norb> <na> <nb>|<opts>" ...]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
bash scripts/sweep.sh ${TAG}_sweep "$@" > /dev/null 2>&1
python - <<PY
import json
for l in open('gpurun_out/${TAG}_sweep.jsonl'):
    try: d = json.loads(l)
    except Exception:
        print(l.strip()[:300]); continue
    print(d['opts'] or 'default', '|', d['plan'].split(';')[1][:150], '| both %.2f alpha %.2f beta %.2f' % (d['orbital_rotation_ms'], d['alpha_only_ms'], d['beta_only_ms']))
PY
