set -x
mkdir -p gpurun_out/r02/san gpurun_out/r02/ncu
# sanitizer
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02/san/memcheck.log python tools/sanitize_kernels.py > gpurun_out/r02/san/memcheck.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02/san/racecheck.log python tools/sanitize_kernels.py > gpurun_out/r02/san/racecheck.out 2>&1
tail -3 gpurun_out/r02/san/memcheck.log gpurun_out/r02/san/racecheck.log gpurun_out/r02/san/memcheck.out
