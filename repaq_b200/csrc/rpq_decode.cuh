#pragma once
#include "rpq_common.cuh"
namespace rpq {
__global__ void k_dec_streams() {}
}
