python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 30 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r17_bench.json'))
print('value %.1fM q/s'%(d['value']/1e6), 'ms/step %.3f'%d['ms_per_step'], 'e2e %.1fM (%.3f ms)'%(d['e2e']['value']/1e6, d['e2e']['ms_per_step']), 'frac %.3f kernel_ms %.3f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
