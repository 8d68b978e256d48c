for d in 0 8 24 4; do echo "== CRAB_SK_DEBUG=$d"; CRAB_SK_DEBUG=$d timeout 120 python tools/bench_pdl_ab.py 2>&1 | grep -E "qkv|gateup|down| o "; done
