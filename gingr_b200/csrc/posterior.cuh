// posterior.cuh -- host-side interface of the K3 posterior kernels (gram.cu, chol.cu, vecops.cu).
#pragma once
#include "common.cuh"

namespace gingr {
}  // namespace gingr
