# split-K factor per decode linear, judged by the in-kernel timeline (last exit to last exit per launch)
for cfg in "" "o:8" "o:2" "qkv:4" "qkv:1" "d:8" "d:2" "gu:2" ; do
  echo "SPLITS=$cfg"
  CRAB_SKINNY_SPLITS="$cfg" timeout 200 python tools/trace_skinny.py --tag _sweep 2>&1 | grep "last-exit to\|traced step" | sed 's/in-kernel timeline.*replay, //'
done
