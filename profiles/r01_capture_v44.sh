# v44: factory.make_woodpile + Crystal._fourier_fields (host-side additions). GPU parity suite + smoke.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v44.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v44.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke_v44.log 2>&1; tail -1 gpurun_out/r01_smoke_v44.log
