# A/B of compile-time tuning variants (build/libemdee_<name>.so built with -D...; selected with EMDEE_B200_LIB).
# usage: bash tools/gpu_variants.sh [name ...]     (the in-tree library runs first as "base")
# build a variant here first, e.g.
#   mkdir -p build && (cd emdee.jl_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
#     -Xcompiler -fPIC -shared -DFLP_PRELOAD2=1 -o ../../build/libemdee_preload2.so emdee_b200.cu -ldl)
# and check parity of a variant with  EMDEE_B200_LIB=/root/repo/build/libemdee_<name>.so python -m pytest tests -m gpu -q
cd /root/repo
B="timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1"
run() { # name, env...
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'bricks' gpurun_out/s_$name.err | head -1 | sed 's/.emdee. //;s/(full.*//') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.4f kernel %.4f build %.3f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch']))" 2>&1 | tail -1)"
}
run base $B
for v in "$@"; do run $v EMDEE_B200_LIB=/root/repo/build/libemdee_$v.so $B; done
