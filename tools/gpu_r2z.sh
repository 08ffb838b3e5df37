# after the model-curve mode of the per-star kernel: smoke + the whole GPU suite (bench kernels are byte-identical in SASS)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v "Missing units" | tail -2
python -m pytest tests -m gpu -q 2>&1 | grep -v "Missing units" | tail -15 | tee gpurun_out/r2z_pytest.log
