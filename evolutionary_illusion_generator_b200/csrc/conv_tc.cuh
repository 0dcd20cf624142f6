// tcgen05 tensor-core 3x3 convolution (3xTF32) - placeholder until the kernel lands; reports "unavailable"
// so eig_set_conv_mode(EIG_CONV_TC) fails loudly instead of silently using another path.
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"
#include <string>

namespace eig {
struct TcWeights { void* d = nullptr; };
inline bool tc_available() { return false; }
inline std::string tc_unavailable_reason() { return "tcgen05 convolution kernel not built yet"; }
inline std::string tc_last_error() { return "tcgen05 convolution kernel not built yet"; }
inline void tc_free(TcWeights&) {}
inline int tc_pack(TcWeights&, const float*, int, int, int) { return 0; }
inline int tc_conv(const TcWeights&, const ConvArgs&, cudaStream_t) { return -1; }
}  // namespace eig
