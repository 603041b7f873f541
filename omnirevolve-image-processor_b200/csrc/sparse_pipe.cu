// sparse_pipe.cu -- placeholder, filled in below
#include "fast_device.cuh"
