# compute-sanitizer is closed on this GPU pool; instead the GPU tests run against a library built with -DEMDEE_CHECKS=1
# (explicit bounds checks on the staged indices, per-lane stacks / rows and task indices of the list kernels; a violation
# raises device flag 8 and the next getter fails).  Build first:
#   (cd emdee.jl_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
#      -DEMDEE_CHECKS=1 -o ../../build/libemdee_checked.so emdee_b200.cu -ldl)
cd /root/repo
mkdir -p gpurun_out
EMDEE_B200_LIB=/root/repo/build/libemdee_checked.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_checked.log; cat gpurun_out/tests_checked.log
EMDEE_B200_LIB=/root/repo/build/libemdee_checked.so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_checked.json 2> gpurun_out/bench_checked.err; tail -1 gpurun_out/bench_checked.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_checked.json') if l.startswith('{')][-1]); print('checked build: c3 ms/step %.4f parity %s'%(d['ms_per_step'], d['parity']['ok']))"
