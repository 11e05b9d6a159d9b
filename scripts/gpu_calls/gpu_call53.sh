#!/bin/bash
# final validation of the round: whole GPU suite, smoke, the default bench line, sanitizer over every kernel
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2/c53_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2/c53_bench.json 2> gpurun_out/r2/c53_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2/c53_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/c53_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cfg4', d.get('cfg4',{}).get('value'), 'cfg5', d.get('cfg5',{}).get('value'))
print('attention', json.dumps(d.get('attention_read'))[:900])
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r2/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2/sanitizer_racecheck.log
exit 0
