#!/bin/bash
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh TAG [workloads...]
# runs the GPU test-suite, then one bench line per workload into gpurun_out/TAG_<workload>.json
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_gpu.txt
tail -3 gpurun_out/${TAG}_gpu.txt
for w in "$@"; do
  python bench.py --workload $w > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err || tail -5 gpurun_out/${TAG}_$w.err
  python tools/show_bench.py gpurun_out/${TAG}_$w.json ${TAG}_$w
done
