#!/bin/bash
# e2e (host buffers) of the default bench against the number of channel groups of a host call: bash tools/e2e_groups.sh 4 8 16 32
for g in "$@"; do
    NEO_B200_HOST_GROUPS=$g python bench.py --no-modes --no-fft-sweep --steps 10 2>/dev/null > /tmp/e2e_$g.json
    python - "$g" /tmp/e2e_$g.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print("groups", sys.argv[1], "e2e", round(d["e2e"]["value"]), "value", round(d["value"]))
PY
done
