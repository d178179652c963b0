// quick standalone compile of the n = 8 first-generation kernels (registers / spills without rebuilding the whole library)
#include "h3d_kernels.cuh"
using namespace h3d;
void* fns[] = {(void*)k_volume<8, 0, true, false, true>, (void*)k_volume<8, 0, true, false, false>, (void*)k_gradient<8, true, false, true>, (void*)k_gradient<8, true, false, false>};
