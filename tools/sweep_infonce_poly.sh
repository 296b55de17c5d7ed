# Offload ratio of the InfoNCE exponentials (POLY of every 8 on the FMA pipe): build one library per ratio (CPU, nvcc) with
#   bash tools/sweep_infonce_poly.sh build
# and time them on the GPU with
#   bash tools/sweep_infonce_poly.sh run
set -e
cd "$(dirname "$0")/.."
mkdir -p build/poly gpurun_out/r03
if [ "$1" = "build" ]; then
  for P in 3 4 5; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
      -DGCF_POLY_OF_8=$P -I include -c recommendation_b200/csrc/infonce.cu -o build/poly/infonce_p$P.o
    OBJS=$(ls build/gcf/*.o | grep -v "/infonce.o")
    /usr/local/cuda/bin/nvcc -shared -o build/poly/libgcf_poly$P.so $OBJS build/poly/infonce_p$P.o -gencode arch=compute_100a,code=sm_100a -cudart static -lcublas
    echo "built build/poly/libgcf_poly$P.so"
  done
else
  for P in 2 3 4 5; do
    L=build/poly/libgcf_poly$P.so; [ $P = 2 ] && L=recommendation_b200/libgcf.so
    echo "== POLY = $P of 8 ($L)"
    GCF_LIB_PATH=$PWD/$L python tools/bench_infonce.py "ssl_layer items" | python -c "import sys,json; [print('  %-26s fwd %.4f ms (%.0f TFLOP/s)  bwd %.4f ms' % (j['case'], j['fwd_ms'], j['fwd_tflops'], j['bwd_ms'])) for j in map(json.loads, sys.stdin)]"
    GCF_LIB_PATH=$PWD/$L python tools/bench_infonce.py "gcl users" | python -c "import sys,json; [print('  %-26s fwd %.4f ms (%.0f TFLOP/s)  bwd %.4f ms' % (j['case'], j['fwd_ms'], j['fwd_tflops'], j['bwd_ms'])) for j in map(json.loads, sys.stdin)]"
    GCF_LIB_PATH=$PWD/$L python -m pytest tests/test_gpu_infonce.py -q -x 2>&1 | tail -1
  done
fi
