// pshadow_fft.cuh -- table-twiddle 4096-point transform (forward: dataset spectra in psh_fft_prepare;
// both directions: test hook), included by pshadow.cu.  The scan's own inverse transform (packed fp32
// arithmetic) and the data formats live in pshadow_fft2.cuh.
//
// FFT: N = 4096 = 16 x 16 x 16, 256 threads, 16 complex values per thread, three radix-16
// passes in registers with two padded shared-memory exchanges.  Input index n = tid + 256 i,
// output index k = tid + 256 c: both coalesced, no bit-reversal pass.
#pragma once

namespace fftx {

constexpr int N = 4096;
constexpr int THREADS = 256;
constexpr int EX_STRIDE = 257;                    // padded row of the exchange buffer (float2)
constexpr int EX_FLOAT2 = 16 * EX_STRIDE;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by exp(DIR * i * angle) given w = exp(+i * angle)
template <int DIR>
__device__ __forceinline__ float2 cmul_dir(float2 a, float2 w) {
    return DIR > 0 ? cmul(a, w) : cmul(a, make_float2(w.x, -w.y));
}

// 4-point DFT, in place, natural order; DIR = -1 forward (e^{-2 pi i nk/4}), +1 inverse
template <int DIR>
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    const float2 it3 = DIR > 0 ? make_float2(-t3.y, t3.x) : make_float2(t3.y, -t3.x);  // (DIR i) t3
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, it3);
    a3 = csub(t1, it3);
}

// 16-point DFT of v[0..15] (natural order in, natural order out), all indices compile-time
template <int DIR>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
    // step 1: for each n1, 4-point DFT over n2 of x[n1 + 4 n2] -> y[n1][k2] kept at v[n1 + 4 k2]
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) fft4<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // step 2: twiddles w16^(n1 k2)
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
    const float d = (float)DIR;
    v[1 + 4 * 1] = cmul(v[1 + 4 * 1], make_float2(c1, d * s1));    // w^1
    v[2 + 4 * 1] = cmul(v[2 + 4 * 1], make_float2(r, d * r));      // w^2
    v[3 + 4 * 1] = cmul(v[3 + 4 * 1], make_float2(s1, d * c1));    // w^3
    v[1 + 4 * 2] = cmul(v[1 + 4 * 2], make_float2(r, d * r));      // w^2
    {                                                              // w^4 = DIR i
        const float2 t = v[2 + 4 * 2];
        v[2 + 4 * 2] = DIR > 0 ? make_float2(-t.y, t.x) : make_float2(t.y, -t.x);
    }
    v[3 + 4 * 2] = cmul(v[3 + 4 * 2], make_float2(-r, d * r));     // w^6
    v[1 + 4 * 3] = cmul(v[1 + 4 * 3], make_float2(s1, d * c1));    // w^3
    v[2 + 4 * 3] = cmul(v[2 + 4 * 3], make_float2(-r, d * r));     // w^6
    v[3 + 4 * 3] = cmul(v[3 + 4 * 3], make_float2(-c1, -d * s1));  // w^9
    // step 3: for each k2, 4-point DFT over n1 -> X[4 k1 + k2] lands at v[k1 + 4 k2]
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) fft4<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    // natural order: X[4 k1 + k2] = v[k1 + 4 k2]  (a register renaming under full unrolling)
    float2 w[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) w[4 * k1 + k2] = v[k1 + 4 * k2];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = w[i];
}

// powers w^1 .. w^15 of a unit complex number from w^1 and w^4 (both table-exact): every power
// is a product of at most three table values, so its error stays below ~17 u (a table value: u)
__device__ __forceinline__ void unit_powers(float2 w1, float2 w4, float2 (&w)[16]) {
    w[1] = w1;
    w[2] = cmul(w1, w1);
    w[3] = cmul(w[2], w1);
    w[4] = w4;
    w[5] = cmul(w4, w1);
    w[6] = cmul(w4, w[2]);
    w[7] = cmul(w4, w[3]);
    w[8] = cmul(w4, w4);
    w[9] = cmul(w[8], w1);
    w[10] = cmul(w[8], w[2]);
    w[11] = cmul(w[8], w[3]);
    w[12] = cmul(w[8], w4);
    w[13] = cmul(w[12], w1);
    w[14] = cmul(w[12], w[2]);
    w[15] = cmul(w[12], w[3]);
}

// loop-invariant twiddle seeds of a thread: pass A uses powers of exp(2 pi i tid/4096), pass B
// powers of exp(2 pi i (tid&15)/256)
struct TwSeeds { float2 a1, a4, b1, b4; };
__device__ __forceinline__ TwSeeds load_seeds(const float2 *__restrict__ tw, int tid) {
    TwSeeds s;
    const int t1 = tid & 15;
    s.a1 = __ldg(tw + tid);
    s.a4 = __ldg(tw + 4 * tid);
    s.b1 = __ldg(tw + 16 * t1);
    s.b4 = __ldg(tw + 64 * t1);
    return s;
}

// 4096-point transform by one CTA of 256 threads.  In: v[i] = x[tid + 256 i].  Out: v[c] =
// X[tid + 256 c].  tw[m] = exp(+2 pi i m / 4096).  `ex` is EX_FLOAT2 float2 of shared memory;
// the caller must __syncthreads() before ex is touched again.
// TABLE = true : every inter-pass twiddle is read from the table (accuracy: dataset spectra)
// TABLE = false: twiddles are rebuilt from four register-resident seeds (no loads in the loop)
template <int DIR, bool TABLE>
__device__ __forceinline__ void fft4096(float2 (&v)[16], float2 *ex, const float2 *__restrict__ tw, int tid,
                                        const TwSeeds &seeds) {
    fft16<DIR>(v);  // over i -> a
    if (TABLE) {
#pragma unroll
        for (int a = 1; a < 16; ++a) v[a] = cmul_dir<DIR>(v[a], __ldg(tw + a * tid));
    } else {
        float2 w[16];
        unit_powers(seeds.a1, seeds.a4, w);
#pragma unroll
        for (int a = 1; a < 16; ++a) v[a] = cmul_dir<DIR>(v[a], w[a]);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) ex[a * EX_STRIDE + tid] = v[a];
    __syncthreads();
    const int a2 = tid >> 4, t1 = tid & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ex[a2 * EX_STRIDE + t1 + 16 * i];
    fft16<DIR>(v);  // over tau2 -> b
    if (TABLE) {
#pragma unroll
        for (int b = 1; b < 16; ++b) v[b] = cmul_dir<DIR>(v[b], __ldg(tw + 16 * b * t1));
    } else {
        float2 w[16];
        unit_powers(seeds.b1, seeds.b4, w);
#pragma unroll
        for (int b = 1; b < 16; ++b) v[b] = cmul_dir<DIR>(v[b], w[b]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 16; ++b) ex[a2 * EX_STRIDE + b * 16 + t1] = v[b];
    __syncthreads();
    const int a = tid & 15, bb = tid >> 4;
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = ex[a * EX_STRIDE + bb * 16 + t];
    fft16<DIR>(v);  // over tau1 -> c
}


}  // namespace fftx
