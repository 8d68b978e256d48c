for mb in 0 16 32 64 96; do
  echo "PREFETCH_MB=$mb" >> gpurun_out/r2_pf_sweep.log
  CRAB_PREFETCH_MB=$mb timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs "" 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['rooflines']
print('  decode_step_ms', round(j['phases']['decode_step_ms'],3), {k[7:]:round(v['us_per_step']/ (32 if k!='decode_lm_head' else 1),1) for k,v in r.items() if k.startswith('decode_')})" >> gpurun_out/r2_pf_sweep.log 2>&1
done
