# Lane-kernel configuration sweep (consumer warps x resident CTAs built into build/lib_cw*_mb*.so, ring depth by SPG_STAGES); run under gpurun.
mkdir -p gpurun_out
run() { timeout 200 python bench.py --no-e2e --no-cpu --steps 5 --pool 4 2>> gpurun_out/r2h_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), d['roofline']['kernel_ms_mean'])"; }
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_per or full_length or synthetic" 2>&1 | tail -3
unset SPG_LIB; for st in 0 4 5 6 8; do export SPG_STAGES=$st; run "cw12_mb2_st$st"; done
unset SPG_STAGES
for v in cw10_mb2 cw14_mb2 cw8_mb3 cw9_mb3; do export SPG_LIB=$PWD/build/lib_$v.so; run $v; done
