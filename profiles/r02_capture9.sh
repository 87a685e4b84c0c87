# Round 2, capture 9: GPU tests after the bail-out fix, host profile of the scalar loop
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c9.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c9.log
python profiles/scalar_loop_profile.py > gpurun_out/r02_scalar_profile_c9.txt 2>&1; head -60 gpurun_out/r02_scalar_profile_c9.txt
