// K2 placeholder (filled in next): warp-per-QP kernel for 8 < n <= 32.
#include "ldp_device.cuh"
bool ldp_warp_supports(const LdpLayout &) { return false; }
int launch_ldp_warp(const LdpLayout &, const LdpSettings &, const double *, const ThetaSegs &, int64_t,
                    const SolveOut &, const SimDesc *, int, void *) { return (int)cudaErrorNotSupported; }
