# last check of the committed tree on a fresh box: build check is a no-op (prebuilt .so travel), smoke, GPU tests, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | grep -v "Missing units" | tail -3
python -m pytest tests -m gpu -q 2>&1 | grep -v "Missing units" | tail -3
python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; python tools/bench_digest.py gpurun_out/r02s_bench.json | head -2
