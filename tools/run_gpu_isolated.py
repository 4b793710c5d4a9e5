"""Runs every gpu-marked test in its own process (a trapped kernel poisons the CUDA context of the process
that launched it), with a per-test timeout.  Debug helper for the GPU box: python tools/run_gpu_isolated.py [paths...]"""
import subprocess
import sys
import time

paths = sys.argv[1:] or ['tests']
ids = subprocess.run([sys.executable, '-m', 'pytest', '--collect-only', '-q', '-m', 'gpu'] + paths, capture_output=True, text=True).stdout
ids = [l.strip() for l in ids.splitlines() if '::' in l]
print(f'{len(ids)} tests', flush=True)
fails = 0
for t in ids:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', t], capture_output=True, text=True, timeout=300)
        ok = r.returncode == 0
        tail = '' if ok else '\n'.join((r.stdout + r.stderr).splitlines()[-25:])
    except subprocess.TimeoutExpired:
        ok, tail = False, 'TIMEOUT'
    fails += not ok
    print(f'{"PASS" if ok else "FAIL"} {t} ({time.time() - t0:.1f}s)', flush=True)
    if not ok:
        print(tail, flush=True)
print(f'failures: {fails}/{len(ids)}')
sys.exit(1 if fails else 0)
