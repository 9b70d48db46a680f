F="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-render"
for args in "--bg-chunks 2" "--bg-chunks 2 --bg-stages 6" "--bg-chunks 3" "--bg-chunks 2 --bg-stages 8"; do
  echo "== $args"
  python bench.py $F $args 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], json.dumps(d['phase_ms']))"
done
