# A/B of the persistent software-pipelined fused frame kernel (default) against the one-CTA-per-unit form
for knob in "" "NEO_B200_FRAME_NO_PIPELINE=1"; do
  echo "== [$knob] T=256 Q=4"; env $knob timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
  echo "== [$knob] T=256 Q=2"; env $knob FRAME_TAPS=524288 timeout 120 python tools/frame_time.py 256 2>&1 | tail -1
  echo "== [$knob] T=512 Q=2"; env $knob timeout 120 python tools/frame_time.py 512 2>&1 | tail -1
  echo "== [$knob] T=128 Q=8"; env $knob timeout 120 python tools/frame_time.py 128 2>&1 | tail -1
done
