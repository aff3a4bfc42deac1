for cfg in "8 4" "8 2" "12 2" "16 2" "16 4" "12 4"; do set -- $cfg; python bench.py --steps 20 --warmup 3 --no-cpu --no-big --no-nn --e2e-depth $1 --e2e-lanes $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth $1 lanes $2 e2e', d['e2e']['value'], 'value', d['value'])"; done
