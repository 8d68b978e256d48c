# split-K factor per decode linear, judged by the in-kernel timeline (last exit to last exit per launch)
for cfg in "" "d:4" "d:2" "o:2" "o:8" "qkv:4" "qkv:1" "gu:2" "d:4,o:2" ; do
  echo "SPLITS=$cfg"
  CRAB_SKINNY_SPLITS="$cfg" timeout 200 python tools/trace_skinny.py --tag _sweep 2>&1 | grep "last-exit to\|traced step" | sed 's/in-kernel timeline.*replay, //'
done
