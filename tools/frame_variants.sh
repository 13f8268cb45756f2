for v in 0 1 2 3 4; do
  echo "== variant $v T=512"; NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 512 2>&1 | tail -1
  echo "== variant $v T=256 Q=2"; FRAME_TAPS=524288 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
done
echo "== variant 1 T=256 Q=4"; NEO_B200_FRAME_VARIANT=1 timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
