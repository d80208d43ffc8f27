# build/libemdee_<name>.so from the tree's sources with extra -D flags:  bash tools/build_variant.sh <name> [-D...]
# (selected at run time with EMDEE_B200_LIB=/root/repo/build/libemdee_<name>.so; see tools/gpu_variants.sh)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
(cd emdee.jl_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
    -o ../../build/libemdee_$name.so emdee_b200.cu -ldl)
echo "built build/libemdee_$name.so ($*)"
