#!/bin/bash
# A/B of environment knobs on the FFT sweep: bash tools/fft_ab.sh "X=1" "NEO_B200_C2R_TWO_CTAS=1" ...
for e in "$@"; do
    env $e python bench.py --fft-only 2>/dev/null > /tmp/fft_ab.json
    python - "$e" <<'PY'
import json, sys
d = json.load(open("/tmp/fft_ab.json"))["fft_sweep"]
print(sys.argv[1], [(r["n"], round(r["r2c_frac"], 3), round(r["c2r_frac"], 3)) for r in d])
PY
done
