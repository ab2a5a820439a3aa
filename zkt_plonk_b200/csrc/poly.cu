// poly.cu -- grand-product, quotient and polynomial utility kernels (filled in below as the path widens).
#include "ctx.h"
#include "ff.cuh"
