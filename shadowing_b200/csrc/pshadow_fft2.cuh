// pshadow_fft2.cuh -- FFT flavour of the filter scan (PSH_MODE_FFT), round 2.  Included by pshadow.cu.
//
// The cross term D_t = sum_j q_j y_{t+j} of ||q - y_t||^2 = Q2 + Y2_t - 2 D_t is a correlation: for a
// whole trajectory it costs O(log N) per window through the FFT instead of W FMAs.  Two trajectories
// share one complex transform (z = y_a + i y_b; q is real, so corr(z, q) = corr(y_a, q) + i corr(y_b, q)).
// What is query-independent is computed once per dataset (psh_fft_prepare):
//   * Zh  : the 4096-point spectra of all row pairs, quantised to fp16 pairs with a per-pair power-of-two
//           scale (4 bytes per complex point);  the quantisation error ||Z^ - Z||_2 / sqrt(N) is MEASURED
//           per pair (zqerr) and enters the rigorous slack through Cauchy-Schwarz:
//           |D^_t - D_t| <= (1/N) sum_k |dZ_k| |Q_k| <= ||dZ||_2 ||q||_2 / sqrt(N);
//   * Y2h : the window energies Y2[r][t], scaled by a per-pair power of two and rounded DOWN to fp16
//           (a lower bound stays a lower bound; the upper bound gives the 2^-10 back), the two rows of a
//           pair interleaved in the scan's output order (4 bytes per pair of windows);
//   * per pair: ||y_pair|| (rounded up), zqerr, the energy scale es and m2 = -2 es / zs.
// Together 8 bytes per pair of samples = exactly the bytes of the raw rows: the scan streams the
// ALGORITHMIC bytes (round 1 streamed fp32 spectra + fp32 energies, 2.05 x).
//
// Per query the scan multiplies by conj(FFT(q))/N, runs ONE inverse 4096-point FFT per row pair in
// registers / shared memory and tests, in the pair's scaled units,
//     fma(m2, v, yf) <= rhs,   rhs = (thr - (Q2 - slack)) * es,
// i.e. the rigorous lower bound LB = Q2 + Y2^ - 2 D^ - slack <= thr of round 1; survivors go through
// the exact re-rank, so results stay bit-identical to the exact scan.
//
// Arithmetic: Blackwell's packed fp32 pipe (add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  A
// complex value lives in one 64-bit register pair; a complex add is ONE instruction, a multiplication
// by +-i folds into an FFMA2 with a swapped / half-negated operand, a complex multiplication is two
// (FMUL2 with a broadcast operand + FFMA2): the radix-16 butterfly issues 81 instructions instead of
// 160.  Each lane is an IEEE fp32 operation, so the error analysis of round 1 is unchanged.
//
// FFT: N = 4096 = 16 x 16 x 16, 256 threads, 16 complex values per thread, three radix-16 passes in
// registers; exchange 1 through padded shared memory (the transform's only CTA barrier), exchange 2 is
// a 16 x 16 transpose inside each half-warp.  Input index n = tid + 256 i; output v[c] = X[k],
// k = (tid >> 4) + 16 (tid & 15) + 256 c (the energies are stored in that order).
// (Measured and rejected this round: four radix-8 passes with 512 threads -- 64 registers per thread, 32
// resident warps per SM -- spilled ~200 bytes per thread and added a third exchange: 0.33 ms against
// 0.23 ms; pass-B twiddles from a shared table and 3 CTAs per SM at 80 registers were slower as well.)
#pragma once
// (pshadow.cu includes <cuda_fp16.h> at file scope)

namespace fx2 {

constexpr int N = 4096;
constexpr int THREADS = 256;
constexpr int EX_STRIDE = 272;                    // 16 x 17 float2 per row: conflict-free transposes
constexpr int EX_FLOAT2 = 16 * EX_STRIDE;

#ifndef PSH_SCALAR_FFT
#define PSH_PK2(name, op)                                                                              \
    __device__ __forceinline__ float2 name(float2 a, float2 b) {                                       \
        float2 r;                                                                                      \
        asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; " op                     \
            " rc, ra, rb; mov.b64 {%0,%1}, rc;}"                                                       \
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                          \
        return r;                                                                                      \
    }
PSH_PK2(add2, "add.rn.f32x2")
PSH_PK2(sub2, "sub.rn.f32x2")
PSH_PK2(mul2, "mul.rn.f32x2")
#undef PSH_PK2
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#else   // A/B build: the same transform on the scalar fp32 pipe (FADD / FMUL / FFMA)
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
// a + i b and a - i b: one FFMA2 each (operand b swapped, one lane negated)
__device__ __forceinline__ float2 add_i(float2 a, float2 b) { return fma2(make_float2(b.y, b.x), make_float2(-1.0f, 1.0f), a); }
__device__ __forceinline__ float2 sub_i(float2 a, float2 b) { return fma2(make_float2(b.y, b.x), make_float2(1.0f, -1.0f), a); }
// a * w = a.x (w.x, w.y) + a.y (-w.y, w.x): FMUL2 (broadcast a.y, swapped w) + FFMA2 (broadcast a.x, addend half-negated)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 p = mul2(make_float2(a.y, a.y), make_float2(w.y, w.x));
    return fma2(make_float2(a.x, a.x), w, make_float2(-p.x, p.y));
}

// inverse 4-point DFT (e^{+2 pi i nk/4}), in place, natural order
__device__ __forceinline__ void ifft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3), t3 = sub2(a1, a3);
    a0 = add2(t0, t2);
    a2 = sub2(t0, t2);
    a1 = add_i(t1, t3);
    a3 = sub_i(t1, t3);
}
// the same with a2 standing for i * a2 (the w16^4 twiddle folded into the butterfly)
__device__ __forceinline__ void ifft4_a2i(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = add_i(a0, a2), t1 = sub_i(a0, a2), t2 = add2(a1, a3), t3 = sub2(a1, a3);
    a0 = add2(t0, t2);
    a2 = sub2(t0, t2);
    a1 = add_i(t1, t3);
    a3 = sub_i(t1, t3);
}

// inverse 16-point DFT of v[0..15], natural order in and out; 81 packed instructions
__device__ __forceinline__ void ifft16(float2 (&v)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) ifft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
    v[1 + 4 * 1] = cmul(v[1 + 4 * 1], make_float2(c1, s1));     // w^1
    v[2 + 4 * 1] = cmul(v[2 + 4 * 1], make_float2(r, r));       // w^2
    v[3 + 4 * 1] = cmul(v[3 + 4 * 1], make_float2(s1, c1));     // w^3
    v[1 + 4 * 2] = cmul(v[1 + 4 * 2], make_float2(r, r));       // w^2
    //  v[2 + 4 * 2] *= w^4 = i : folded into the k2 = 2 butterfly below
    v[3 + 4 * 2] = cmul(v[3 + 4 * 2], make_float2(-r, r));      // w^6
    v[1 + 4 * 3] = cmul(v[1 + 4 * 3], make_float2(s1, c1));     // w^3
    v[2 + 4 * 3] = cmul(v[2 + 4 * 3], make_float2(-r, r));      // w^6
    v[3 + 4 * 3] = cmul(v[3 + 4 * 3], make_float2(-c1, -s1));   // w^9
    ifft4(v[0], v[1], v[2], v[3]);
    ifft4(v[4], v[5], v[6], v[7]);
    ifft4_a2i(v[8], v[9], v[10], v[11]);
    ifft4(v[12], v[13], v[14], v[15]);
    // natural order: X[4 k1 + k2] = v[k1 + 4 k2]  (a register renaming under full unrolling)
    float2 w[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) w[4 * k1 + k2] = v[k1 + 4 * k2];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = w[i];
}

// v[a] *= w^a, a = 1..15, the powers built from w^1 and w^4 (both table-exact): every power is a
// product of at most three table values (error <= ~17 u, inside CF).  Interleaved so that only
// w1, w2, w3 and one of w4 / w8 / w12 are live at a time.
__device__ __forceinline__ void twiddle_powers(float2 (&v)[16], float2 w1, float2 w4) {
    const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1)); v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
    const float2 w8 = cmul(w4, w4);
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
    const float2 w12 = cmul(w8, w4);
    v[12] = cmul(v[12], w12);
    v[13] = cmul(v[13], cmul(w12, w1)); v[14] = cmul(v[14], cmul(w12, w2)); v[15] = cmul(v[15], cmul(w12, w3));
}

// loop-invariant twiddle seeds of a thread: pass A uses powers of exp(2 pi i tid/4096), pass B powers of
// exp(2 pi i (tid&15)/256)
struct Seeds { float2 a1, a4, b1, b4; };
__device__ __forceinline__ Seeds load_seeds(const float2 *__restrict__ tw, int tid) {
    Seeds s;
    const int t1 = tid & 15;
    s.a1 = __ldg(tw + tid);
    s.a4 = __ldg(tw + 4 * tid);
    s.b1 = __ldg(tw + 16 * t1);
    s.b4 = __ldg(tw + 64 * t1);
    return s;
}

// inverse 4096-point transform by one CTA of 256 threads.  In: v[i] = x[tid + 256 i].
// Out: v[c] = X[(tid >> 4) + 16 (tid & 15) + 256 c].  `after_first_barrier()` runs right behind the
// transform's only CTA barrier: every thread has consumed the staged spectrum by then;
// `before_first_barrier()` right in front of it (what it writes to shared memory is visible behind it).
template <typename F0, typename F>
__device__ __forceinline__ void ifft4096(float2 (&v)[16], float2 *ex, int tid, const Seeds &seeds,
                                         F0 before_first_barrier, F after_first_barrier) {
    ifft16(v);  // over i -> a
    twiddle_powers(v, seeds.a1, seeds.a4);
#pragma unroll
    for (int a = 0; a < 16; ++a) ex[a * EX_STRIDE + tid] = v[a];
    before_first_barrier();
    __syncthreads();
    after_first_barrier();
    const int a2 = tid >> 4, t1 = tid & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ex[a2 * EX_STRIDE + t1 + 16 * i];
    ifft16(v);  // over tau2 -> b
    twiddle_powers(v, seeds.b1, seeds.b4);
    // row a2 of ex is read (above) and rewritten (below) by the 16 threads of this half-warp only
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 16; ++b) ex[a2 * EX_STRIDE + b * 17 + t1] = v[b];
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = ex[a2 * EX_STRIDE + t1 * 17 + t];
    ifft16(v);  // over tau1 -> c
}

// window of output register c of thread tid: kb(tid) + 256 c
__device__ __forceinline__ int out_base(int tid) { return (tid >> 4) + 16 * (tid & 15); }

}  // namespace fx2
