// quick standalone compile of the second-generation kernels (registers / spills without rebuilding the whole library):
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xptxas -v -fmad=false -I include -I horses3d_b200/csrc -cubin -o /tmp/k2.cubin scripts/dev/k2.cu
#include "h3d_kernels2.cuh"
using namespace h3d;
void* fns[] = {(void*)k_volume2<false>, (void*)k_volume2<true>, (void*)k_gradient2<false>, (void*)k_gradient2<true>};
