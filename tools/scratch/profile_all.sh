bash tools/scratch/profile.sh
bash tools/scratch/profile2.sh
timeout 300 python tools/timeline.py > gpurun_out/prof/timeline.log 2>&1; echo "timeline exit $?"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/prof/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/prof/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/prof/smoke.log 2>&1; tail -1 gpurun_out/prof/smoke.log
for a in "--points 1000000 --cloud torus --k 256 --order 3 --queries 1000000" "--points 10000000 --cloud multi --k 256 --order 3 --queries 1000000" "--points 10000000 --cloud multi --k 512 --order 4 --queries 262144" "--points 30000000 --cloud density --k 256 --order 3 --queries 1000000"; do timeout 300 python tools/scale_probe.py $a 2>/dev/null | tail -1 >> gpurun_out/prof/scale_probe.jsonl; done; cat gpurun_out/prof/scale_probe.jsonl | cut -c1-250
timeout 300 python tools/scratch/k512_probe.py 512 4 2000000 2>/dev/null | tail -1 > gpurun_out/prof/k512.jsonl; cat gpurun_out/prof/k512.jsonl
