# Round-2 evidence for profiles/ (run under gpurun): launch list of a short bench run, one ncu --set full capture of the headline kernel
# (exported as raw / source CSV: the .ncu-rep itself is too large to bring back), the same for configs 3 and 4.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --pairs-per-step 2000000 --pool 2 --configs C3,C4,C5 --config-steps 2 --config-pool 2 --parity-pairs 0 --no-e2e --no-cpu > gpurun_out/launches_r2.log 2>&1
for c in C2 C3 C4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:trim_lanes -s 3 -c 1 -o /tmp/r2_$c python bench.py --only $c --steps 3 --pairs-per-step 2000000 --pool 2 --parity-pairs 0 > gpurun_out/ncu_r2_$c.log 2>&1
  ncu -i /tmp/r2_$c.ncu-rep --page raw --csv > gpurun_out/ncu_r2_${c}_raw.csv
  ncu -i /tmp/r2_$c.ncu-rep --page source --csv | gzip > gpurun_out/ncu_r2_${c}_src.csv.gz
done
ls -la gpurun_out | tail -12
