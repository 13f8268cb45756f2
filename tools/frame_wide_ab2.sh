# where the 32-points-per-thread fused frame kernel wins: variant 0 (16 points per thread) against 5, by T and Q
for v in 0 5; do
  echo "== variant $v T=512 Q=1"; FRAME_TAPS=524288 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 512 2>&1 | tail -1
  echo "== variant $v T=512 Q=4"; FRAME_TAPS=2097152 FRAME_CHANNELS=512 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 512 2>&1 | tail -1
  echo "== variant $v T=256 Q=3"; FRAME_TAPS=786432 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
  echo "== variant $v T=256 Q=1"; FRAME_TAPS=262144 NEO_B200_FRAME_VARIANT=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
done
