# Lane-kernel configuration sweep (consumer warps x resident CTAs built into build/lib_*.so with -DSPG_LANE_CW[_LONG] / -DSPG_LANE_MINB[_LONG],
# ring depth by SPG_STAGES); run under gpurun:  bash profiles/lane_config_sweep.sh C3
mkdir -p gpurun_out
cfg=${1:-C2}
run() { timeout 200 python bench.py --only $cfg --steps 5 --pool 3 --parity-pairs 0 2>> gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$cfg', round(d['value'],1), d['kernel'])"; }
unset SPG_LIB; run main
for f in build/lib_*.so; do export SPG_LIB=$PWD/$f; run $(basename $f .so); done
