# final validation of a tree on one B200: the whole GPU suite, smoke(), the default bench line
tag=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -x -q -rf > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 200 gpurun_out/${tag}_bench.json; echo
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value']); print(d['surface'])"
