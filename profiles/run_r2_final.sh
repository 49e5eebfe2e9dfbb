#!/bin/bash
# round 2, final verification on one B200: the whole GPU suite, smoke(), the command-line comparison, the default bench line
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -4 gpurun_out/final_smoke.log
SPG_CLI_TOOLS=0 SPG_CLI_ONLY=seqpurge_b200,b200_plain_in_level0,b200_bgzf python profiles/cli_throughput.py 4000000 > gpurun_out/cli_throughput_r2.log 2>&1; cat gpurun_out/cli_throughput_r2.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['e2e']['value'], d['e2e'].get('parity'), d['roofline']['frac'], {k:(v['value']) for k,v in d['configs'].items()}, d['parity'], d['cpu_baseline']['value'])"
