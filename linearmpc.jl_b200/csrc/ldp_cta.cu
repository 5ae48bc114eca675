// K3 placeholder (filled in next): CTA-per-QP kernel for n > 32.
#include "ldp_device.cuh"
bool ldp_cta_supports(const LdpLayout &) { return false; }
int launch_ldp_cta(const LdpLayout &, const LdpSettings &, const double *, const ThetaSegs &, int64_t,
                   const SolveOut &, const SimDesc *, int, void *) { return (int)cudaErrorNotSupported; }
