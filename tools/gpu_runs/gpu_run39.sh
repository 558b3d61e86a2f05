#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 0 1; do echo "== PDL=$pdl"; LDM_PDL=$pdl python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'])"; done
