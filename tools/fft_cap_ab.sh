# A/B on one box: 64-register caps for the 2^10..2^12-point real-transform kernels (default build) vs the 72/80/85/128 caps
# (NEO_B200_LIBRARY=neo-dsp_b200/libneo_b200_cap80.so): BASELINE config 2 sweep and the convolver's r2c / c2r phases
for lib in "" "NEO_B200_LIBRARY=$PWD/neo-dsp_b200/libneo_b200_cap80.so"; do
  echo "== [$lib] fft sweep"; env $lib timeout 120 python bench.py --fft-only 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(' '.join(f\"{r['n']}:{r['r2c_frac']:.2f}/{r['c2r_frac']:.2f}\" for r in d['fft_sweep']))"
  echo "== [$lib] frame 256"; env $lib timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
done
