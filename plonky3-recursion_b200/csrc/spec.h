// spec.h — registry of build-time specialised quotient kernels (defined in spec.cu from specialized_gen.cuh).
#pragma once
#include "kernels.cuh"
namespace p3r {
typedef void (*SpecQuotientKernel)(QuotientArgs);
struct SpecEntry {
    uint64_t hash;      // FNV-1a over the Montgomery-encoded instruction words of the constraint program
    int field_id;
    uint32_t n_insns;
    SpecQuotientKernel fn;
};
typedef void (*SpecLogupKernel)(LogupArgs);
struct SpecLogupEntry {
    uint64_t hash;      // FNV-1a over the lookup-input program words and the lookup / interaction structure
    int field_id;
    uint32_t n_insns;
    SpecLogupKernel fn;
};
const SpecLogupEntry* p3r_spec_logup_registry(size_t* n);
unsigned p3r_spec_logup_threads();
void p3r_spec_logup_launch(SpecLogupKernel fn, const LogupArgs& a, unsigned grid, cudaStream_t stream);
const SpecEntry* p3r_spec_registry(size_t* n);
unsigned p3r_spec_threads();   // threads per CTA of the generated kernels (32 rows x constraint groups)
void p3r_spec_launch(SpecQuotientKernel fn, const QuotientArgs& a, unsigned grid, unsigned block, cudaStream_t stream);
}  // namespace p3r
