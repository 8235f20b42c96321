// spec.cu — separate translation unit for the generated straight-line quotient kernels (slow to compile, rarely changes).
#include "spec.h"
#include "specialized_gen.cuh"
namespace p3r {
const SpecEntry* p3r_spec_registry(size_t* n) {
    *n = sizeof(SPEC_QUOTIENT) / sizeof(SPEC_QUOTIENT[0]);
    return SPEC_QUOTIENT;
}
unsigned p3r_spec_threads() { return SPEC_THREADS; }
const SpecLogupEntry* p3r_spec_logup_registry(size_t* n) {
    *n = sizeof(SPEC_LOGUP) / sizeof(SPEC_LOGUP[0]);
    return SPEC_LOGUP;
}
unsigned p3r_spec_logup_threads() { return SPEC_LOGUP_THREADS; }
void p3r_spec_logup_launch(SpecLogupKernel fn, const LogupArgs& a, unsigned grid, cudaStream_t stream) {
    fn<<<grid, SPEC_LOGUP_THREADS, 0, stream>>>(a);
}
void p3r_spec_launch(SpecQuotientKernel fn, const QuotientArgs& a, unsigned grid, unsigned block, cudaStream_t stream) {
    fn<<<grid, block, 0, stream>>>(a);
}
}  // namespace p3r
