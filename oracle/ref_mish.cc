// ref_mish.cc — compiles the REFERENCE's own Mish math (mmdet/ops/mish_cuda/src/mish.h:17-29, included from
// /root/reference where it lies, never copied) into oracle/_ref/libmish_ref.so, so that the restatement in oracle.c
// (oracle_mish_fwd / oracle_mish_bwd) and the CUDA kernels can be checked against the real thing.
// TEST INFRASTRUCTURE: only tests/ load the result. Built by `make -C oracle ref` (needs /root/reference and the
// installed torch headers, which mish.h includes; the result links against libm only).
#include REF_MISH_H

extern "C" {
void ref_mish_fwd(const float* in, float* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = mish_fwd_func<float>(in[i]);  // what mish_cpu.cc:8-15 runs per element
}
void ref_mish_bwd(const float* grad_out, const float* in, float* grad_in, long long n) {
    for (long long i = 0; i < n; ++i) grad_in[i] = mish_bwd_func<float>(grad_out[i], in[i]);  // mish_cpu.cc:19-28
}
}
