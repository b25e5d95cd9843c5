// Host-side shim (ours) around the reference's kernel.h: exposes its hash()
// (kernel.h:157-180) and the RandomEngine it aliases (thrust taus88, kernel.h:20)
// plus thrust's uniform_real_distribution<float>, so tests can pin the C
// restatement in evogp_oracle.c against the real thing.  Test infrastructure.
#include "kernel.h"
#include <cstdint>

extern "C" {
uint32_t ref_hash(uint32_t n, uint32_t k1, uint32_t k2) { return hash(n, k1, k2); }
void ref_engine_draws(uint32_t seed, int n, uint32_t* out) {
    RandomEngine e(seed);
    for (int i = 0; i < n; i++) out[i] = e();
}
void ref_engine_uniforms(uint32_t seed, int n, float* out) {
    RandomEngine e(seed);
    thrust::uniform_real_distribution<float> u(0.0f, 1.0f);
    for (int i = 0; i < n; i++) out[i] = u(e);
}
uint32_t ref_default_engine_nth(int n) {
    RandomEngine e;
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v = e();
    return v;
}
}
