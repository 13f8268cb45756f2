# A/B of the 32-points-per-thread fused frame kernel (NEO_B200_FRAME_VARIANT=5) against the shipped forms
for v in 0 5; do
  echo "== variant $v T=256 Q=4"; NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
  echo "== variant $v T=256 Q=2"; FRAME_TAPS=524288 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
  echo "== variant $v T=512 Q=2"; NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 512 2>&1 | tail -1
done
