#!/bin/bash
# round 2, third session: parity of the stream kernels after the batched segment copies, per-kernel stream profile
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_fastq.py tests/test_gpu_readqc.py tests/test_gpu_fastqtrim.py tests/test_gpu_cli.py -x -q -m gpu > gpurun_out/r2s_tests.log 2>&1
tail -3 gpurun_out/r2s_tests.log
python profiles/fq_kernels.py 1000000 100000 > gpurun_out/r2s_fq_100k.txt 2>&1; cp gpurun_out/fq_kernels.json gpurun_out/r2s_fq_100k.json
python profiles/fq_kernels.py 2000000 1000000 > gpurun_out/r2s_fq_1m.txt 2>&1; cp gpurun_out/fq_kernels.json gpurun_out/r2s_fq_1m.json
cat gpurun_out/r2s_fq_100k.txt gpurun_out/r2s_fq_1m.txt
