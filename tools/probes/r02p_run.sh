# final check of the last build (N fold): full GPU test suite, smoke(), bench line
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
tail -2 gpurun_out/r02p_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02p_bench.json'))
print(d['value'], d['e2e']['value'], d['config']['batch64']['crops_per_s'], d['clocks'])
"
