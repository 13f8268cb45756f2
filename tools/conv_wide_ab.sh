# A/B of the convolver's block transforms at B = 1024 (C5, frame mode T = 256): r2c through cta_fft (default) or the wide kernel compiled for
# 4 / 5 / 6 resident CTAs (NEO_B200_CONV_WIDE_R2C), c2r through the wide kernel at 4 / 5 / 6 (NEO_B200_CONV_WIDE_C2R)
echo "== default"; timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
for v in 4 5 6; do
  echo "== r2c wide $v, c2r wide $v"; NEO_B200_CONV_WIDE_R2C=$v NEO_B200_CONV_WIDE_C2R=$v timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
done
