python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/s4_bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ctc_roofline']['frac'], d['ctc_roofline']['ms_per_call'], d['roofline']['frac'], d['gpu_launches'])"
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
