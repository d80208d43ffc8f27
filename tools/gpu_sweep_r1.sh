cd /root/repo
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-iters 1"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'force kernel mode' gpurun_out/s_$name.err | head -3 | sed 's/.*mode//;s/.emdee. //' | tr '\n' ';')"
}
run exp_noz $B --dt 1e-9 --temperature 0.0001
run exp_ref EMDEE_FUSE=0 $B --dt 1e-9 --temperature 0.0001
