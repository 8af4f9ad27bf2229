// C-ABI glue: error text, version queries, mask decoding shared by the kernels.
#include <stdarg.h>
#include <stdio.h>

#include "cnf_common.cuh"

namespace cnf {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int build_mask(const cnf_mask& m, int C, MaskView* out) {
    MaskView v{};
    if (C < 1 || C > CNF_MAX_CHANNELS) return fail(CNF_ERR_UNSUPPORTED, "C=%d outside [1, %d]", C, CNF_MAX_CHANNELS);
    for (int c = 0; c < C; ++c) {
        const bool cond = m.cond_c_host != nullptr && m.cond_c_host[c] != 0.0f;
        if (cond) v.cond_c |= (1ull << c);
        else v.tch[v.n_t++] = (unsigned char)c;
    }
    v.c0 = v.n_t > 0 ? v.tch[0] : 0;
    v.contiguous = 1;
    for (int j = 1; j < v.n_t; ++j)
        if (v.tch[j] != v.tch[j - 1] + 1) v.contiguous = 0;
    if (m.cond_s_host != nullptr && m.s_period > 0) {
        if (m.s_period > 64) return fail(CNF_ERR_UNSUPPORTED, "position mask period %d > 64", m.s_period);
        v.s_period = m.s_period;
        for (int s = 0; s < m.s_period; ++s)
            if (m.cond_s_host[s] != 0.0f) v.cond_s |= (1ull << s);
        if (v.cond_s == 0) v.s_period = 0;
    }
    *out = v;
    return CNF_OK;
}

int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

}  // namespace cnf

extern "C" const char* cnf_last_error_string(void) { return cnf::g_err; }
extern "C" int cnf_abi_version(void) { return 5; }
extern "C" int cnf_built_for_sm(void) { return 100; }
