#!/bin/bash
# usage (GPU box, repo root): bash tools/gpu_final.sh TAG   -- the round's final single-GPU measurements into gpurun_out/
T=$1
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${T}_gpu_tests.txt; cat $O/${T}_gpu_tests.txt
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err || tail -5 $O/${T}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err || tail -5 $O/${T}_ref.err
for w in car-512 pidray-256-labelmap kmeans-assign; do
  python bench.py --workload $w > $O/${T}_bench_$w.json 2> $O/${T}_$w.err || tail -5 $O/${T}_$w.err
done
# launch list of the bench command (after it exited 0 without ncu above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_raw.csv \
  python bench.py --steps 2 --warmup 1 --no-alt --no-eager --no-cpu-baseline > $O/${T}_ncu_bench.log 2>&1
python tools/summarize_launches.py $O/${T}_launches_raw.csv $O/${T}_launches.csv "$T: python bench.py --steps 2 --warmup 1 --no-alt --no-eager --no-cpu-baseline" || true
rm -f $O/${T}_launches_raw.csv
# full captures of the kernels changed in this session: the Sinkhorn pass that writes the 16-bit cache (launch 24 of the
# probe: the last of its 12 writer launches) and the first pass over the cache (launch 25)
python tools/gpu_probe_sk16.py 160000 5000 > $O/${T}_sk16_probe.txt 2>&1 && ncu --set full --clock-control none --import-source on \
  -k regex:sinkhorn_pass -s 24 -c 2 -o $O/${T}_sk16 python tools/gpu_probe_sk16.py 160000 5000 > $O/${T}_ncu_sk16.log 2>&1
python tools/gpu_probe_sk16.py > $O/${T}_sk16_accuracy.txt 2>&1
python tools/gpu_probe_loss.py > $O/${T}_loss_probe.txt 2>&1
for f in $O/${T}_bench_n1.json $O/${T}_bench_car-512.json $O/${T}_bench_pidray-256-labelmap.json $O/${T}_bench_kmeans-assign.json; do python tools/show_bench.py $f x | head -3; done
ls -la $O | tail -30
